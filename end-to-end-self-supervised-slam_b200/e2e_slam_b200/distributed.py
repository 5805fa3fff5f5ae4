"""Data parallelism over independent key-frame pairs.

The hot path has no exchange step: pairs (and source frames, and scales) are independent units, so each
rank runs the fused kernels on its own shard with no data-path collective.  The only collective of a
refinement step is the all-reduce of the depth network's adaptation gradients (about 57 MB of fp32,
SURVEY.md section 5), which follows the network's backward pass -- not one of our kernels -- so there is
nothing to fuse it into; it is issued as ONE flat bucket on a side stream so that it overlaps whatever the
main stream does next.  PointFusion over one sequence does not shard (frame s fuses into the map of frame
s-1): replicas only, one sequence per rank.

One process per GPU (torchrun); backend "nccl" on GPUs, "gloo" in the CPU tests.

On an NVSwitch box the bucket's all-reduce is OUR kernel (csrc/multimem_allreduce.cu, `FlatGradBucket(..., nvls=True)`): the bucket
is allocated in symmetric memory, rank r reduces slice r inside the switch with `multimem.ld_reduce` and broadcasts the mean with
`multimem.st`, bracketed by two cross-rank barriers -- a few CTAs for the time the switch needs, instead of NCCL's ring kernel holding
16-32 SMs next to an issue-bound sweep.  torch.distributed._symmetric_memory provides allocation, rendezvous and the barrier
(plumbing); without multicast support the bucket falls back to the NCCL all-reduce.
"""
import ctypes
import os

import torch
import torch.distributed as dist


def shard_pairs(n_pairs, rank, world_size):
    """Indices of the pairs rank `rank` owns: rank::world_size (round robin keeps shards within one pair of
    each other for any n_pairs)."""
    if not (0 <= rank < world_size):
        raise ValueError(f"rank {rank} outside world of size {world_size}")
    return list(range(rank, n_pairs, world_size))


class FlatGradBucket:
    """All-reduce(mean) of the gradients of `params` as one flat fp32 bucket.

    Parameters that never receive a gradient (the reference's depth net has 8: resnet `fc` and the unused
    `dispconv` scales) must be skipped *identically on every rank*, so membership
    is decided by `requires_grad` at construction time, not by `grad is None` at reduce time; a member whose
    grad is None on some rank contributes zeros.  Frozen parameters (refinement mode freezes every tensor
    whose name contains "bn", train_depth.py:213-222) are not members."""

    def __init__(self, params, process_group=None, device=None, nvls=False, nvls_ctas=0):
        self.params = [p for p in params if p.requires_grad]
        if not self.params:
            raise ValueError("no trainable parameters")
        self.group = process_group
        self.device = torch.device(device) if device is not None else self.params[0].device
        self.numel = sum(p.numel() for p in self.params)
        self.stream = torch.cuda.Stream(device=self.device) if self.device.type == "cuda" else None
        self._work = None
        self._symm, self.nvls_ctas, self.nvls_error = None, int(nvls_ctas), None
        self.flat = None
        if nvls and self.device.type == "cuda" and os.environ.get("E2E_NVLS", "1") != "0":
            self._try_nvls()
        if self.flat is None:
            self.flat = torch.zeros(self.numel, dtype=torch.float32, device=self.device)

    def _try_nvls(self):
        """Bucket in symmetric memory with a multicast mapping; any failure (no NVSwitch, no multicast support, an older torch)
        leaves the NCCL path in place and records why."""
        try:
            import torch.distributed._symmetric_memory as symm_mem
            group = self.group if self.group is not None else dist.group.WORLD
            padded = (self.numel + 3) // 4 * 4
            buf = symm_mem.empty(padded, dtype=torch.float32, device=self.device)
            hdl = symm_mem.rendezvous(buf, group)
            if not int(getattr(hdl, "multicast_ptr", 0)):
                raise RuntimeError("the symmetric-memory handle has no multicast address (no NVLS on this system)")
            buf.zero_()
            self._symm, self._symm_buf, self._padded = hdl, buf, padded
            self.flat = buf[:self.numel]
        except Exception as e:                        # noqa: BLE001 -- every reason means "use NCCL"
            self._symm, self.flat = None, None
            self.nvls_error = f"{type(e).__name__}: {e}"[:300]

    @property
    def uses_nvls(self):
        return self._symm is not None

    def _views(self):
        off = 0
        for p in self.params:
            yield p, self.flat[off:off + p.numel()].view_as(p)
            off += p.numel()

    def adopt_grads(self):
        """Make every member's `.grad` a view INTO the flat bucket (what DDP calls gradient_as_bucket_view): backward then
        accumulates straight into the bucket and start() / finish() move no bytes except the all-reduce itself.  Existing
        gradients are copied in once."""
        for p, v in self._views():
            if p.grad is not None:
                v.copy_(p.grad)
            else:
                v.zero_()
            p.grad = v
        return self

    def _is_view(self, p, v):
        return p.grad is not None and p.grad.data_ptr() == v.data_ptr() and p.grad.shape == v.shape

    def start(self):
        """Launch the all-reduce(mean) (asynchronously on the side stream on GPUs); gradients that are not bucket views are
        packed first.  NCCL averages inside the collective (ReduceOp.AVG); gloo sums and the division follows."""
        world = dist.get_world_size(self.group)
        if self.stream is not None:
            self.stream.wait_stream(torch.cuda.current_stream(self.device))
            ctx = torch.cuda.stream(self.stream)
        else:
            ctx = torch.autograd.profiler.record_function("e2e.bucket")
        with ctx:
            for p, v in self._views():
                if self._is_view(p, v):
                    continue
                if p.grad is None:
                    v.zero_()
                else:
                    v.copy_(p.grad)
            if self._symm is not None:
                from ._lib import check, lib
                hdl = self._symm
                hdl.barrier(channel=0)                  # every rank's gradients are in its copy of the bucket
                check(lib().e2e_multimem_allreduce_avg(ctypes.c_void_p(int(hdl.multicast_ptr)), self._padded, hdl.rank, hdl.world_size,
                                                       self.nvls_ctas, ctypes.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)),
                      "e2e_multimem_allreduce_avg")
                hdl.barrier(channel=1)                  # every slice is reduced and written back everywhere
                self._work = True
            elif self.device.type == "cuda":
                self._work = dist.all_reduce(self.flat, op=dist.ReduceOp.AVG, group=self.group, async_op=True)
            else:
                self.flat.div_(world)
                self._work = dist.all_reduce(self.flat, op=dist.ReduceOp.SUM, group=self.group, async_op=True)
        return self

    def finish(self):
        """Wait for the all-reduce and hand the averaged gradients back (no copy for gradients that are bucket views)."""
        if self._work is None:
            raise RuntimeError("finish() called before start()")
        if self._work is not True:
            self._work.wait()
        if self.stream is not None:
            torch.cuda.current_stream(self.device).wait_stream(self.stream)
        for p, v in self._views():
            if self._is_view(p, v):
                continue
            if p.grad is None:
                p.grad = v.clone()
            else:
                p.grad.copy_(v)
        self._work = None

    def allreduce(self):
        self.start()
        self.finish()


def mean_over_ranks(value, group=None):
    """Average a scalar tensor over ranks (the logged loss); a 4-byte all-reduce."""
    v = value.detach().clone().reshape(1)
    dist.all_reduce(v, op=dist.ReduceOp.SUM, group=group)
    return v / dist.get_world_size(group)
