// Exact K = 1 nearest neighbour through a uniform grid: the same answer as the brute-force kernel of knn.cu (and as
// chamferdist.chamfer.knn_points, loss/losses.py:39-63; online_adaption.py:638-645), bit for bit -- squared distances in the
// oracle's operation order ((dx*dx + dy*dy) + dz*dz), lowest index among exact ties -- at a cost that does not grow with
// P1 * P2.  The point-supervision loss of the online loop queries 307 200 live points against a map of millions
// ("dominant cost of the online loop once the map is large"): 6e11 pairs by brute force, ~1e8 here.
//
//   bbox      min / max of the finite reference points (ordered-integer atomics)
//   params    one thread: cell size for ~4 points per cell, at most 2^21 cells; all grid parameters stay on the device
//   count     cell of every reference point, histogram
//   scan      exclusive prefix sum of the histogram (one CTA)
//   fill      reference points sorted by cell as {x, y, z, index} records
//   query     one thread per query: cells at Chebyshev distance 0, 1, 2, ... around the query's cell until the best
//             distance found is provably smaller than anything outside the searched cube (or the cube covers the grid)
// Nothing synchronises with the host.
#include "common.cuh"

namespace e2e {

constexpr int KG_NT = 256;
constexpr int KG_MAX_CELLS = 1 << 21;

struct GridParams {
    float ox, oy, oz;      // origin (bbox minimum)
    float h, inv_h;        // cell size
    int nx, ny, nz;        // cells per axis
    int ncells;
};

__device__ __forceinline__ unsigned ordered_bits(float f)      // monotone float -> unsigned
{
    const unsigned u = __float_as_uint(f);
    return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__device__ __forceinline__ float from_ordered(unsigned o)
{
    return __uint_as_float((o & 0x80000000u) ? (o & 0x7fffffffu) : ~o);
}

__global__ void __launch_bounds__(KG_NT) kg_bbox_kernel(const float *ref, long long P2, unsigned *bb)
{
    float lo[3] = {INFINITY, INFINITY, INFINITY}, hi[3] = {-INFINITY, -INFINITY, -INFINITY};
    for (long long i = (long long)blockIdx.x * KG_NT + threadIdx.x; i < P2; i += (long long)gridDim.x * KG_NT) {
        const float x = ref[i * 3], y = ref[i * 3 + 1], z = ref[i * 3 + 2];
        if (isfinite(x) && isfinite(y) && isfinite(z)) {
            lo[0] = fminf(lo[0], x); lo[1] = fminf(lo[1], y); lo[2] = fminf(lo[2], z);
            hi[0] = fmaxf(hi[0], x); hi[1] = fmaxf(hi[1], y); hi[2] = fmaxf(hi[2], z);
        }
    }
#pragma unroll
    for (int k = 0; k < 3; k++) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            lo[k] = fminf(lo[k], __shfl_xor_sync(0xffffffffu, lo[k], o));
            hi[k] = fmaxf(hi[k], __shfl_xor_sync(0xffffffffu, hi[k], o));
        }
        if ((threadIdx.x & 31) == 0) {
            atomicMin(bb + k, ordered_bits(lo[k]));
            atomicMax(bb + 3 + k, ordered_bits(hi[k]));
        }
    }
}

__global__ void kg_params_kernel(const unsigned *bb, long long P2, GridParams *gp)
{
    float lo[3], ext[3];
    for (int k = 0; k < 3; k++) {
        lo[k] = from_ordered(bb[k]);
        const float hi = from_ordered(bb[3 + k]);
        ext[k] = (hi >= lo[k]) ? hi - lo[k] : 0.0f;        // no finite point at all: one cell at the origin
        if (!(hi >= lo[k])) lo[k] = 0.0f;
    }
    const float emax = fmaxf(fmaxf(ext[0], ext[1]), fmaxf(ext[2], 1e-20f));
    double target = (double)P2 / 4.0;
    if (target < 1.0) target = 1.0;
    if (target > (double)KG_MAX_CELLS) target = (double)KG_MAX_CELLS;
    // cell size from the volume of the box (thin extents count as one cell), then enlarged until the grid fits
    double vol = 1.0;
    for (int k = 0; k < 3; k++) vol *= fmax((double)ext[k], (double)emax * 1e-3);
    float h = (float)cbrt(vol / target);
    if (!(h > emax * 1e-6f)) h = emax * 1e-6f;
    int n[3];
    for (int it = 0; it < 64; it++) {
        double cells = 1.0;
        for (int k = 0; k < 3; k++) {
            n[k] = (int)fmin(floor((double)ext[k] / (double)h) + 1.0, 2097152.0);
            cells *= (double)n[k];
        }
        if (cells <= (double)KG_MAX_CELLS) break;
        h *= 1.26f;
    }
    gp->ox = lo[0]; gp->oy = lo[1]; gp->oz = lo[2];
    gp->h = h; gp->inv_h = 1.0f / h;
    gp->nx = n[0]; gp->ny = n[1]; gp->nz = n[2];
    gp->ncells = n[0] * n[1] * n[2];
}

// cell coordinate along one axis; anything not representable (NaN, huge) goes far outside on a definite side
__device__ __forceinline__ int cell_coord(float v, float o, float inv_h)
{
    const float c = floorf((v - o) * inv_h);
    if (!(c > -1.0e6f)) return -1000000;       // also NaN
    if (c > 1.0e6f) return 1000000;
    return (int)c;
}

__device__ __forceinline__ int ref_cell(const GridParams &g, float x, float y, float z)
{
    if (!(isfinite(x) && isfinite(y) && isfinite(z))) return 0;     // never the nearest of anything: where it sits is irrelevant
    const int cx = min(max(cell_coord(x, g.ox, g.inv_h), 0), g.nx - 1);
    const int cy = min(max(cell_coord(y, g.oy, g.inv_h), 0), g.ny - 1);
    const int cz = min(max(cell_coord(z, g.oz, g.inv_h), 0), g.nz - 1);
    return (cz * g.ny + cy) * g.nx + cx;
}

__global__ void __launch_bounds__(KG_NT) kg_count_kernel(const float *ref, long long P2, const GridParams *gp, int *cell_of, int *count)
{
    const GridParams g = *gp;
    for (long long i = (long long)blockIdx.x * KG_NT + threadIdx.x; i < P2; i += (long long)gridDim.x * KG_NT) {
        const int c = ref_cell(g, ref[i * 3], ref[i * 3 + 1], ref[i * 3 + 2]);
        cell_of[i] = c;
        atomicAdd(count + c, 1);
    }
}

// start[c] = exclusive prefix sum of count (in place in `start`, which holds the counts on entry), cursor = copy
__global__ void __launch_bounds__(1024) kg_scan_kernel(int *start, int *cursor, const GridParams *gp)
{
    __shared__ int wtot[32];
    __shared__ int carry;
    const int n = gp->ncells;
    if (threadIdx.x == 0) carry = 0;
    __syncthreads();
    for (int base = 0; base < n; base += 1024) {
        const int i = base + threadIdx.x;
        const int v = (i < n) ? start[i] : 0;
        int incl = v;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int t = __shfl_up_sync(0xffffffffu, incl, o);
            if ((threadIdx.x & 31) >= o) incl += t;
        }
        if ((threadIdx.x & 31) == 31) wtot[threadIdx.x >> 5] = incl;
        __syncthreads();
        if (threadIdx.x < 32) {
            int w = wtot[threadIdx.x], wi = w;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const int t = __shfl_up_sync(0xffffffffu, wi, o);
                if (threadIdx.x >= o) wi += t;
            }
            wtot[threadIdx.x] = wi - w;
        }
        __syncthreads();
        const int excl = carry + wtot[threadIdx.x >> 5] + incl - v;
        if (i < n) { start[i] = excl; cursor[i] = excl; }
        __syncthreads();
        if (threadIdx.x == 1023) carry = excl + v;
        __syncthreads();
    }
    if (threadIdx.x == 0) start[n] = carry;
}

__global__ void __launch_bounds__(KG_NT) kg_fill_kernel(const float *ref, long long P2, const int *cell_of, int *cursor, float4 *sorted)
{
    for (long long i = (long long)blockIdx.x * KG_NT + threadIdx.x; i < P2; i += (long long)gridDim.x * KG_NT) {
        const int pos = atomicAdd(cursor + cell_of[i], 1);
        sorted[pos] = make_float4(ref[i * 3], ref[i * 3 + 1], ref[i * 3 + 2], __int_as_float((int)i));
    }
}

__global__ void __launch_bounds__(KG_NT) kg_query_kernel(const float *query, const float *T, long long P1, const GridParams *gp,
                                                         const int *start, const float4 *sorted, float *dist2, long long *idx)
{
    const GridParams g = *gp;
    const long long i = (long long)blockIdx.x * KG_NT + threadIdx.x;
    if (i >= P1) return;
    float q[3];
    {
        const float x = query[i * 3], y = query[i * 3 + 1], z = query[i * 3 + 2];
        if (T) {   // R p + t, accumulated left to right (same as knn.cu / the oracle's transform_pointcloud)
#pragma unroll
            for (int r = 0; r < 3; r++)
                q[r] = xadd(xadd(xadd(xmul(T[r * 4], x), xmul(T[r * 4 + 1], y)), xmul(T[r * 4 + 2], z)), T[r * 4 + 3]);
        } else {
            q[0] = x; q[1] = y; q[2] = z;
        }
    }
    float best = INFINITY;
    int bi = 0;
    if (isfinite(q[0]) && isfinite(q[1]) && isfinite(q[2])) {      // otherwise every distance is NaN / inf: brute force answers (inf, 0)
        const int cx = cell_coord(q[0], g.ox, g.inv_h), cy = cell_coord(q[1], g.oy, g.inv_h), cz = cell_coord(q[2], g.oz, g.inv_h);
        // rings before the first one that can touch the grid hold nothing
        int r = max(max(max(-cx, cx - (g.nx - 1)), max(-cy, cy - (g.ny - 1))), max(max(-cz, cz - (g.nz - 1)), 0));
        const float slack = 0.01f + 4e-7f * (float)max(max(g.nx, g.ny), g.nz);      // cells: rounding of the cell coordinates
        auto visit = [&](int x, int y, int z) {
            const int c = (z * g.ny + y) * g.nx + x;
            const int s = start[c], e = start[c + 1];
            for (int j = s; j < e; j++) {
                const float4 p = sorted[j];
                const float dx = xsub(q[0], p.x), dy = xsub(q[1], p.y), dz = xsub(q[2], p.z);
                const float d2 = xadd(xadd(xmul(dx, dx), xmul(dy, dy)), xmul(dz, dz));
                const int pi = __float_as_int(p.w);
                if (d2 < best || (d2 == best && pi < bi)) { best = d2; bi = pi; }          // first minimum in index order
            }
        };
        for (;; r++) {
            const int x0 = max(cx - r, 0), x1 = min(cx + r, g.nx - 1);
            const int y0 = max(cy - r, 0), y1 = min(cy + r, g.ny - 1);
            const int z0 = max(cz - r, 0), z1 = min(cz + r, g.nz - 1);
            for (int z = z0; z <= z1; z++) {
                const bool zface = (z == cz - r) || (z == cz + r);
                for (int y = y0; y <= y1; y++) {
                    if (zface || y == cy - r || y == cy + r) {          // a face of the cube: the whole row belongs to ring r
                        for (int x = x0; x <= x1; x++) visit(x, y, z);
                    } else {                                            // inside: only the two x faces
                        if (cx - r >= 0 && cx - r <= g.nx - 1) visit(cx - r, y, z);
                        if (cx + r >= 0 && cx + r <= g.nx - 1) visit(cx + r, y, z);
                    }
                }
            }
            // every point within Chebyshev cell distance r has been seen, i.e. every point closer than r*h (less the slack);
            // nothing outside can beat `best` once sqrt(best) <= that radius
            const float reach = ((float)r - slack) * g.h;
            if (reach > 0.0f && best <= reach * reach) break;
            if (cx - r <= 0 && cx + r >= g.nx - 1 && cy - r <= 0 && cy + r >= g.ny - 1 && cz - r <= 0 && cz + r >= g.nz - 1) break;
        }
    }
    dist2[i] = best;
    idx[i] = (long long)bi;
}

static size_t kg_a256(size_t n) { return (n + 255) / 256 * 256; }
static int kg_blocks(long long n)
{
    long long b = (n + KG_NT - 1) / KG_NT;
    if (b > kNumSMs * 16) b = kNumSMs * 16;
    return (int)(b < 1 ? 1 : b);
}

}  // namespace e2e

using namespace e2e;

extern "C" {

size_t e2e_knn1_grid_workspace_bytes(long long P2)
{
    if (P2 < 0) P2 = 0;
    return kg_a256(((size_t)KG_MAX_CELLS + 1) * 4) + kg_a256((size_t)KG_MAX_CELLS * 4) + kg_a256((size_t)P2 * 4) + kg_a256((size_t)P2 * 16) +
           kg_a256(sizeof(GridParams)) + 256 + 256;
}

// Build the grid over `ref` into `workspace` (e2e_knn1_grid_workspace_bytes(P2)); the grid stays valid for any number of
// e2e_knn1_grid_query calls against the same reference cloud (ICP queries it 20-40 times).
int e2e_knn1_grid_build(const float *ref, long long P2, void *workspace, size_t workspace_bytes, void *stream)
{
    cudaStream_t st = (cudaStream_t)stream;
    E2E_REQUIRE(ref && P2 > 0, "knn1_grid: empty or null reference cloud (P2=%lld)", P2);
    E2E_REQUIRE(P2 < (1ll << 31), "knn1_grid: too many reference points");
    E2E_REQUIRE(workspace && workspace_bytes >= e2e_knn1_grid_workspace_bytes(P2), "knn1_grid: workspace too small");
    unsigned char *w = (unsigned char *)workspace;
    int *start = (int *)w;              w += kg_a256(((size_t)KG_MAX_CELLS + 1) * 4);
    int *cursor = (int *)w;             w += kg_a256((size_t)KG_MAX_CELLS * 4);
    int *cell_of = (int *)w;            w += kg_a256((size_t)P2 * 4);
    float4 *sorted = (float4 *)w;       w += kg_a256((size_t)P2 * 16);
    GridParams *gp = (GridParams *)w;   w += kg_a256(sizeof(GridParams));
    unsigned *bb = (unsigned *)w;
    // bbox accumulators: minima start at all ones, maxima at zero (ordered encoding); histogram at zero
    if (cudaMemsetAsync(bb, 0xff, 12, st) != cudaSuccess || cudaMemsetAsync(bb + 3, 0x00, 12, st) != cudaSuccess ||
        cudaMemsetAsync(start, 0, ((size_t)KG_MAX_CELLS + 1) * 4, st) != cudaSuccess)
        return finish_launch("knn1_grid: memset");
    const int nb = kg_blocks(P2);
    kg_bbox_kernel<<<nb, KG_NT, 0, st>>>(ref, P2, bb);
    kg_params_kernel<<<1, 1, 0, st>>>(bb, P2, gp);
    kg_count_kernel<<<nb, KG_NT, 0, st>>>(ref, P2, gp, cell_of, start);
    kg_scan_kernel<<<1, 1024, 0, st>>>(start, cursor, gp);
    kg_fill_kernel<<<nb, KG_NT, 0, st>>>(ref, P2, cell_of, cursor, sorted);
    count_launch(5);
    return finish_launch("knn1_grid_build");
}

int e2e_knn1_grid_query(const float *query, const float *transform, long long P1, long long P2,
                        float *dist2, long long *idx, const void *workspace, void *stream)
{
    cudaStream_t st = (cudaStream_t)stream;
    E2E_REQUIRE(query && dist2 && idx && workspace && P1 > 0 && P2 > 0, "knn1_grid_query: bad arguments");
    const unsigned char *w = (const unsigned char *)workspace;
    const int *start = (const int *)w;          w += kg_a256(((size_t)KG_MAX_CELLS + 1) * 4) + kg_a256((size_t)KG_MAX_CELLS * 4) + kg_a256((size_t)P2 * 4);
    const float4 *sorted = (const float4 *)w;   w += kg_a256((size_t)P2 * 16);
    const GridParams *gp = (const GridParams *)w;
    const long long qb = (P1 + KG_NT - 1) / KG_NT;
    E2E_REQUIRE(qb < (1ll << 31), "knn1_grid: too many query points");
    kg_query_kernel<<<(unsigned)qb, KG_NT, 0, st>>>(query, transform, P1, gp, start, sorted, dist2, idx);
    count_launch();
    return finish_launch("knn1_grid_query");
}

int e2e_knn1_grid_fwd(const float *query, const float *transform, const float *ref, long long P1, long long P2,
                      float *dist2, long long *idx, void *workspace, size_t workspace_bytes, void *stream)
{
    E2E_REQUIRE(query && ref && dist2 && idx && P1 > 0 && P2 > 0, "knn1_grid: empty or null point cloud (P1=%lld, P2=%lld)", P1, P2);
    if (int rc = e2e_knn1_grid_build(ref, P2, workspace, workspace_bytes, stream)) return rc;
    return e2e_knn1_grid_query(query, transform, P1, P2, dist2, idx, workspace, stream);
}

}  // extern "C"
