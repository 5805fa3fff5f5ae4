// Exact K = 1 nearest neighbour through a uniform grid: the same answer as the brute-force kernel of knn.cu (and as
// chamferdist.chamfer.knn_points, loss/losses.py:39-63; online_adaption.py:638-645), bit for bit -- squared distances in the
// oracle's operation order ((dx*dx + dy*dy) + dz*dz), lowest index among exact ties -- at a cost that does not grow with
// P1 * P2.  The point-supervision loss of the online loop queries 307 200 live points against a map of millions
// ("dominant cost of the online loop once the map is large"): 6e11 pairs by brute force, ~1e8 here.
//
//   bbox      min / max of the finite reference points (ordered-integer atomics)
//   params    one thread: first cell size from the box volume (~4 points per cell if the cloud filled its box)
//   count     cell of every reference point, histogram, number of OCCUPIED cells
//   refine    maps are surfaces: a sheet inside its box leaves most cells empty and crowds the occupied ones (2 M points of a
//             room at 2^21 cells: ~290 per occupied cell, and a query 10 cm off the surface scans tens of thousands of
//             points).  So the cell is shrunk (occupancy of a surface goes with h^2) and the count repeated, up to four
//             passes, until an occupied cell holds <= 6 points or the grid reaches 2^24 cells
//   scan      exclusive prefix sum of the histogram (per-chunk sums, scan of the sums, per-chunk rescan)
//   fill      reference points sorted by cell as {x, y, z, index} records
//   query     one thread per query: cells at Chebyshev distance 0, 1, 2, ... around the query's cell until the best
//             distance found is provably smaller than anything outside the searched cube (or the cube covers the grid)
// Nothing synchronises with the host.
#include <climits>
#include <cstdlib>

#include "common.cuh"

namespace e2e {

constexpr int KG_NT = 256;
constexpr int KG_MAX_CELLS = 1 << 24;
constexpr int KG_PASSES = 4;
#ifndef KG_M_VAL
#define KG_M_VAL 6
#endif
constexpr int KG_M = KG_M_VAL;               // fine cells per coarse cell and axis
constexpr int KG_M3 = KG_M * KG_M * KG_M;
// Cell numbering: COARSE-MAJOR.  A fine cell (x, y, z) has the id  coarse(x/M, y/M, z/M) * M^3 + ((z%M) * M + y%M) * M + x%M,  so the
// points of one coarse cell are ONE contiguous range of the sorted records (start[cc * M^3] .. start[(cc + 1) * M^3]) and the far
// search walks it as a flat list, 32 points per warp step.  The occupancy bitmap keeps the LINEAR numbering ((z * ny + y) * nx + x):
// the near search fetches an x-row of occupancy bits with one funnel shift.  ncells counts the padded (coarse-major) ids.

struct GridParams {
    float ox, oy, oz;      // origin (bbox minimum)
    float h, inv_h;        // cell size
    int nx, ny, nz;        // cells per axis
    int ncells;
    float ext[3];          // box extents
    int occupied;          // cells hit by the current count pass
    int done;              // the grid of the last count pass is final
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
    }
    // one pair of atomics per axis and CTA (per warp they were most of this kernel's time: ~19 k atomics on each of six words)
    __shared__ float s_lo[3][KG_NT / 32], s_hi[3][KG_NT / 32];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (lane == 0) {
#pragma unroll
        for (int k = 0; k < 3; k++) { s_lo[k][warp] = lo[k]; s_hi[k][warp] = hi[k]; }
    }
    __syncthreads();
    if (threadIdx.x < 3) {
        const int k = threadIdx.x;
        float l = s_lo[k][0], h = s_hi[k][0];
#pragma unroll
        for (int w = 1; w < KG_NT / 32; w++) { l = fminf(l, s_lo[k][w]); h = fmaxf(h, s_hi[k][w]); }
        atomicMin(bb + k, ordered_bits(l));
        atomicMax(bb + 3 + k, ordered_bits(h));
    }
}

// dims for a cell size (enlarged until the grid fits KG_MAX_CELLS)
__device__ void kg_set_cells(GridParams *gp, float h)
{
    int n[3];
    for (int it = 0; it < 64; it++) {
        double cells = 1.0;
        for (int k = 0; k < 3; k++) {
            n[k] = (int)fmin(floor((double)gp->ext[k] / (double)h) + 1.0, 2097152.0);
            cells *= (double)((n[k] + KG_M - 1) / KG_M * KG_M);      // ids are padded to whole coarse cells per axis
        }
        if (cells <= (double)KG_MAX_CELLS) break;
        h *= 1.26f;
    }
    gp->h = h; gp->inv_h = 1.0f / h;
    gp->nx = n[0]; gp->ny = n[1]; gp->nz = n[2];
    gp->ncells = ((n[0] + KG_M - 1) / KG_M) * ((n[1] + KG_M - 1) / KG_M) * ((n[2] + KG_M - 1) / KG_M) * KG_M3;
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
    // first cell size from the volume of the box (thin extents count as one cell)
    double vol = 1.0;
    for (int k = 0; k < 3; k++) vol *= fmax((double)ext[k], (double)emax * 1e-3);
    float h = (float)cbrt(vol / target);
    if (!(h > emax * 1e-6f)) h = emax * 1e-6f;
    gp->ox = lo[0]; gp->oy = lo[1]; gp->oz = lo[2];
    gp->ext[0] = ext[0]; gp->ext[1] = ext[1]; gp->ext[2] = ext[2];
    gp->occupied = 0;
    gp->done = 0;
    kg_set_cells(gp, h);
}

// after a count pass: fine enough (<= 6 points per occupied cell), or out of passes / cells -> this grid is final;
// otherwise shrink the cell (the occupancy of a surface scales with h^2) and count again
__global__ void kg_refine_kernel(GridParams *gp, long long P2, int last)
{
    if (gp->done) return;
    const float avg = (float)P2 / (float)max(gp->occupied, 1);
    if (avg <= 6.0f || last || gp->ncells > KG_MAX_CELLS / 2) { gp->done = 1; return; }
    float f = sqrtf(4.0f / avg);
    if (f < 0.25f) f = 0.25f;
    gp->occupied = 0;
    kg_set_cells(gp, gp->h * f);
}

__global__ void __launch_bounds__(KG_NT) kg_clear_kernel(int4 *count, int n4, const GridParams *gp)
{
    if (gp->done) return;
    for (int i = blockIdx.x * KG_NT + threadIdx.x; i < n4; i += gridDim.x * KG_NT) count[i] = make_int4(0, 0, 0, 0);
}

// cell coordinate along one axis; anything not representable (NaN, huge) goes far outside on a definite side
__device__ __forceinline__ int cell_coord(float v, float o, float inv_h)
{
    const float c = floorf((v - o) * inv_h);
    if (!(c > -1.0e6f)) return -1000000;       // also NaN
    if (c > 1.0e6f) return 1000000;
    return (int)c;
}

// coarse-major id of the fine cell (x, y, z) (inside the grid)
__device__ __forceinline__ int cell_id(const GridParams &g, int x, int y, int z)
{
    const int mx = (g.nx + KG_M - 1) / KG_M, my = (g.ny + KG_M - 1) / KG_M;
    return (((z / KG_M) * my + y / KG_M) * mx + x / KG_M) * KG_M3 + ((z % KG_M) * KG_M + y % KG_M) * KG_M + x % KG_M;
}

__device__ __forceinline__ int ref_cell(const GridParams &g, float x, float y, float z)
{
    if (!(isfinite(x) && isfinite(y) && isfinite(z))) return 0;     // never the nearest of anything: where it sits is irrelevant
    const int cx = min(max(cell_coord(x, g.ox, g.inv_h), 0), g.nx - 1);
    const int cy = min(max(cell_coord(y, g.oy, g.inv_h), 0), g.ny - 1);
    const int cz = min(max(cell_coord(z, g.oz, g.inv_h), 0), g.nz - 1);
    return cell_id(g, cx, cy, cz);
}

__global__ void __launch_bounds__(KG_NT) kg_count_kernel(const float *ref, long long P2, GridParams *gp, int *cell_of, int *count)
{
    if (gp->done) return;
    const GridParams g = *gp;
    int claimed = 0;
    for (long long i = (long long)blockIdx.x * KG_NT + threadIdx.x; i < P2; i += (long long)gridDim.x * KG_NT) {
        const int c = ref_cell(g, ref[i * 3], ref[i * 3 + 1], ref[i * 3 + 2]);
        cell_of[i] = c;
        claimed += atomicAdd(count + c, 1) == 0;
    }
    claimed = __reduce_add_sync(0xffffffffu, claimed);
    if ((threadIdx.x & 31) == 0 && claimed) atomicAdd(&gp->occupied, claimed);
}

// exclusive scan of the histogram (length padded to whole 4096-entry chunks): per-chunk sums, scan of the sums (one CTA),
// per-chunk rescan
__global__ void __launch_bounds__(1024) kg_scan1_kernel(const int *v, int *chunk_sum, const GridParams *gp)
{
    __shared__ int sh[32];
    if ((long long)blockIdx.x * 4096 > (long long)gp->ncells) return;
    const int4 a = reinterpret_cast<const int4 *>(v)[(size_t)blockIdx.x * 1024 + threadIdx.x];
    int t = a.x + a.y + a.z + a.w;
    t = __reduce_add_sync(0xffffffffu, t);
    if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = t;
    __syncthreads();
    if (threadIdx.x < 32) {
        t = __reduce_add_sync(0xffffffffu, sh[threadIdx.x]);
        if (threadIdx.x == 0) chunk_sum[blockIdx.x] = t;
    }
}
__global__ void __launch_bounds__(1024) kg_scan2_kernel(int *chunk_sum, const GridParams *gp)
{
    __shared__ int wtot[32];
    __shared__ int carry;
    const int nchunks = gp->ncells / 4096 + 1;
    if (threadIdx.x == 0) carry = 0;
    __syncthreads();
    for (int base = 0; base < nchunks; base += 1024) {
        const int i = base + threadIdx.x;
        const int v = (i < nchunks) ? chunk_sum[i] : 0;
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
        if (i < nchunks) chunk_sum[i] = excl;
        __syncthreads();
        if (threadIdx.x == 1023) carry = excl + v;
        __syncthreads();
    }
}
// start = exclusive prefix sums in place (the histogram entry past the last cell is 0, so start[ncells] = P2), cursor = copy
__global__ void __launch_bounds__(1024) kg_scan3_kernel(int *start, const int *chunk_sum, int *cursor, const GridParams *gp)
{
    __shared__ int wtot[32];
    if ((long long)blockIdx.x * 4096 > (long long)gp->ncells) return;
    const size_t base = (size_t)blockIdx.x * 4096 + (size_t)threadIdx.x * 4;
    const int4 a = *reinterpret_cast<const int4 *>(start + base);
    const int mine = a.x + a.y + a.z + a.w;
    int incl = mine;
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
    int e = chunk_sum[blockIdx.x] + wtot[threadIdx.x >> 5] + incl - mine;
    int4 o4;
    o4.x = e; e += a.x; o4.y = e; e += a.y; o4.z = e; e += a.z; o4.w = e;
    *reinterpret_cast<int4 *>(start + base) = o4;
    *reinterpret_cast<int4 *>(cursor + base) = o4;
}

__global__ void __launch_bounds__(KG_NT) kg_fill_kernel(const float *ref, long long P2, const int *cell_of, int *cursor, float4 *sorted)
{
    for (long long i = (long long)blockIdx.x * KG_NT + threadIdx.x; i < P2; i += (long long)gridDim.x * KG_NT) {
        const int pos = atomicAdd(cursor + cell_of[i], 1);
        sorted[pos] = make_float4(ref[i * 3], ref[i * 3 + 1], ref[i * 3 + 2], __int_as_float((int)i));
    }
}

// measured with the flat-list far search (build + query, 307 200 queries 7 / 15 cm off a 2 M-point surface / config C2 step as a CUDA
// graph): M = 2: 1.94 / 3.76 / 5.79 ms, 3: 1.10 / 1.82 / 2.34, 4: 0.93 / 1.39 / 1.70, 5: 0.88 / 1.22 / 1.37, 6: 0.80 / 1.06 / 1.35,
// 7: 0.83 / 1.07 / 1.44, 8: 0.87 / 1.08 / 1.55, 10: 0.96 / 1.16 / 1.82, 12: 1.04 / 1.26 / 2.17 -- smaller cells lengthen the ring walk,
// larger ones lengthen the point lists.  (Round 1's per-lane fine-cell loops had their optimum at M = 4: 1.63 / - / 1.95 ms.)
#ifndef KG_NEAR_RINGS_VAL
#define KG_NEAR_RINGS_VAL 1
#endif
// rings of fine cells searched directly around the query (one thread per query).  Measured with the flat-list far search (build +
// query, 307 200 queries vs 2 M points, mean distance 0 / 2 / 7 / 15 cm; 19 200 x 75 000): 2 rings 0.44 / 0.76 / 1.13 / 1.98 / 0.26 ms,
// 1 ring 0.45 / 0.54 / 0.92 / 1.81 / 0.19 ms -- a second ring walked by single threads is mostly futile work for queries that go
// to the warp-per-query search anyway
constexpr int KG_NEAR_RINGS = KG_NEAR_RINGS_VAL;
#ifndef KG_FAR_MINB
#define KG_FAR_MINB 6      // latency bound: 48 warps per SM (40 registers, a few outer-loop values spilled) measured faster than 32 or 24
#endif

// coarse occupancy counts and one bit per FINE cell (2 MB for 2^24 cells: the emptiness test of a fine cell -- 98 % of
// the cells around a surface are empty -- then stays in cache instead of fetching two words of the 64 MB offset array)
__global__ void __launch_bounds__(KG_NT) kg_coarse_kernel(long long P2, const GridParams *gp, const int *cell_of, int *coarse, unsigned *bits)
{
    const GridParams g = *gp;
    const int mx = (g.nx + KG_M - 1) / KG_M, my = (g.ny + KG_M - 1) / KG_M;
    for (long long i = (long long)blockIdx.x * KG_NT + threadIdx.x; i < P2; i += (long long)gridDim.x * KG_NT) {
        const int id = cell_of[i];                                   // coarse-major id -> coarse cell, linear fine cell
        const int cc = id / KG_M3, f = id % KG_M3;
        const int X = cc % mx, Y = (cc / mx) % my, Z = cc / (mx * my);
        const int x = X * KG_M + f % KG_M, y = Y * KG_M + (f / KG_M) % KG_M, z = Z * KG_M + f / (KG_M * KG_M);
        const int c = (z * g.ny + y) * g.nx + x;
        atomicAdd(coarse + cc, 1);
        atomicOr(bits + (c >> 5), 1u << (c & 31));
    }
}

// Tight bounding box of the points of every occupied coarse cell (one warp per cell, after the fill).  A map is a surface:
// inside a coarse cell it is a thin sheet, and the distance to the sheet's box is a far better bound than the distance to
// the cell (whose near face can be most of a cell closer than anything in it), so the far search drops most neighbours of
// the nearest cell unsearched.  Floating-point subtraction is monotone, so |q - p| >= the box distance holds per axis in
// fp32 exactly; non-finite points (never anyone's nearest) are left out.
__global__ void __launch_bounds__(KG_NT) kg_coarse_box_kernel(const GridParams *gp, const int *start, const float4 *sorted, const int *coarse,
                                                              float4 *cbox)
{
    const GridParams g = *gp;
    const int lane = threadIdx.x & 31;
    const long long total = g.ncells / KG_M3;
    for (long long cc = (long long)blockIdx.x * (KG_NT / 32) + (threadIdx.x >> 5); cc < total; cc += (long long)gridDim.x * (KG_NT / 32)) {
        if (coarse[cc] == 0) {      // empty: an empty range (the far search reads box and range of every candidate cell, never `coarse`)
            if (lane == 0) cbox[2 * cc] = cbox[2 * cc + 1] = make_float4(0.f, 0.f, 0.f, __int_as_float(0));
            continue;
        }
        float lx = INFINITY, ly = INFINITY, lz = INFINITY, hx = -INFINITY, hy = -INFINITY, hz = -INFINITY;
        for (int j = start[cc * KG_M3] + lane, e = start[(cc + 1) * KG_M3]; j < e; j += 32) {      // the cell's points: one contiguous range
            const float4 p = sorted[j];
            if (isfinite(p.x) && isfinite(p.y) && isfinite(p.z)) {
                lx = fminf(lx, p.x); ly = fminf(ly, p.y); lz = fminf(lz, p.z);
                hx = fmaxf(hx, p.x); hy = fmaxf(hy, p.y); hz = fmaxf(hz, p.z);
            }
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            lx = fminf(lx, __shfl_xor_sync(0xffffffffu, lx, o)); ly = fminf(ly, __shfl_xor_sync(0xffffffffu, ly, o));
            lz = fminf(lz, __shfl_xor_sync(0xffffffffu, lz, o)); hx = fmaxf(hx, __shfl_xor_sync(0xffffffffu, hx, o));
            hy = fmaxf(hy, __shfl_xor_sync(0xffffffffu, hy, o)); hz = fmaxf(hz, __shfl_xor_sync(0xffffffffu, hz, o));
        }
        if (lane == 0) {
            // the w slots carry the cell's range of sorted records, so that the far search gets box and range with the same two loads
            cbox[2 * cc] = make_float4(lx, ly, lz, __int_as_float(start[cc * KG_M3]));
            cbox[2 * cc + 1] = make_float4(hx, hy, hz, __int_as_float(start[(cc + 1) * KG_M3]));
        }
    }
}

// Query, phase 1 -- one thread per query: the fine cells within KG_NEAR_RINGS of the query's cell, ring by ring.  A query
// this close to the cloud (the usual case) is decided here; the others are appended to `far_list` with what they have found.
__global__ void __launch_bounds__(KG_NT) kg_query_near_kernel(const float *query, const float *T, long long P1, const GridParams *gp,
                                                              const int *start, const float4 *sorted, const unsigned *bits,
                                                              float *dist2, long long *idx, int *far_list, int *far_count)
{
    const GridParams g = *gp;
    const long long i = (long long)blockIdx.x * KG_NT + threadIdx.x;
    const bool live = i < P1;
    float qx = 0.f, qy = 0.f, qz = 0.f;
    if (live) {
        const float x = query[i * 3], y = query[i * 3 + 1], z = query[i * 3 + 2];
        if (T) {   // R p + t, accumulated left to right (same as knn.cu / the oracle's transform_pointcloud)
            qx = xadd(xadd(xadd(xmul(T[0], x), xmul(T[1], y)), xmul(T[2], z)), T[3]);
            qy = xadd(xadd(xadd(xmul(T[4], x), xmul(T[5], y)), xmul(T[6], z)), T[7]);
            qz = xadd(xadd(xadd(xmul(T[8], x), xmul(T[9], y)), xmul(T[10], z)), T[11]);
        } else {
            qx = x; qy = y; qz = z;
        }
    }
    float best = INFINITY;
    int bi = 0;
    // a query with a non-finite coordinate: every distance is NaN / inf and brute force answers (inf, 0)
    const bool finite = live && isfinite(qx) && isfinite(qy) && isfinite(qz);
    bool decided = !finite;
    if (finite) {
        const int cx = cell_coord(qx, g.ox, g.inv_h), cy = cell_coord(qy, g.oy, g.inv_h), cz = cell_coord(qz, g.oz, g.inv_h);
        const float slack = 0.01f + 4e-7f * (float)max(max(g.nx, g.ny), g.nz);      // cells: rounding of the cell coordinates
        // The occupancy bits of an x-row of cells are adjacent in the bitmap: one funnel shift fetches the row, and only occupied
        // cells are visited (a loop over all 27 / 125 cells with a test per cell leaves most lanes idle while a few scan points).
        // Rings 0-1 first (the 3 x 3 rows of three cells), then, if the nearest point is not yet proven, ring 2 (5 x 5 rows of
        // five cells less the part already searched).
#pragma unroll 1
        for (int R = 1; R <= KG_NEAR_RINGS && !decided; R++) {
            const int xlo = max(cx - R, 0), xhi = min(cx + R, g.nx - 1);
            if (xlo <= xhi) {
                const unsigned xmask = (2u << (xhi - xlo)) - 1u;                    // xhi - xlo <= 4
                // the cells of this row that rings < R covered: x in [cx - R + 1, cx + R - 1], if the row itself was covered
                const int ilo = max(cx - R + 1, xlo), ihi = min(cx + R - 1, xhi);
                const unsigned inner = (R > 1 && ilo <= ihi) ? (((2u << (ihi - ilo)) - 1u) << (ilo - xlo)) : 0u;
#pragma unroll 1
                for (int dz = -R; dz <= R; dz++) {
                    const int z = cz + dz;
                    if (z < 0 || z >= g.nz) continue;
#pragma unroll 1
                    for (int dy = -R; dy <= R; dy++) {
                        const int y = cy + dy;
                        if (y < 0 || y >= g.ny) continue;
                        const int c0 = (z * g.ny + y) * g.nx + xlo;
                        unsigned m = __funnelshift_r(bits[c0 >> 5], bits[(c0 >> 5) + 1], c0 & 31) & xmask;
                        if (abs(dz) < R && abs(dy) < R) m &= ~inner;
                        const int rowid = cell_id(g, 0, y, z);                  // coarse-major id of the row's cell x = 0
                        while (m) {
                            const int x = xlo + __ffs(m) - 1;
                            const int c = rowid + (x / KG_M) * KG_M3 + x % KG_M;
                            m &= m - 1;
                            for (int j = start[c], e = start[c + 1]; j < e; j++) {
                                const float4 p = sorted[j];
                                const float dx = xsub(qx, p.x), dy2 = xsub(qy, p.y), dz2 = xsub(qz, p.z);
                                const float d2 = xadd(xadd(xmul(dx, dx), xmul(dy2, dy2)), xmul(dz2, dz2));
                                const int pi = __float_as_int(p.w);
                                if (d2 <= best) { if (d2 < best || pi < bi) { best = d2; bi = pi; } }      // first minimum in index order
                            }
                        }
                    }
                }
            }
            // every point within Chebyshev cell distance R has been seen, i.e. every point closer than R*h (less the slack)
            const float reach = ((float)R - slack) * g.h;
            decided = best <= reach * reach;
        }
    }
    if (live) {
        dist2[i] = best;
        idx[i] = (long long)bi;
    }
    const unsigned und = __ballot_sync(0xffffffffu, !decided);       // warp-aggregated append
    if (und) {
        const int lane = threadIdx.x & 31;
        int base = 0;
        if (lane == 0) base = atomicAdd(far_count, __popc(und));
        base = __shfl_sync(0xffffffffu, base, 0);
        if (!decided) far_list[base + __popc(und & ((1u << lane) - 1))] = (int)i;
    }
}

// Query, phase 1 for FEW queries (ICP: 19 200 live points per iteration): one thread per query leaves a B200 nearly empty and
// every thread walks its 27 cells through three dependent loads each (measured 42 us per call, 85 of the 125 us of a GradICP
// iteration).  Here a team of TEAM lanes shares one query: the nine (dz, dy) rows of ring 1 are dealt out to the lanes, the
// lanes' (distance, index) are merged with a lexicographic minimum -- the same answer, ~6x less latency.
template <int TEAM>
__global__ void __launch_bounds__(KG_NT) kg_query_near_team_kernel(const float *query, const float *T, long long P1, const GridParams *gp,
                                                                   const int *start, const float4 *sorted, const unsigned *bits,
                                                                   float *dist2, long long *idx, int *far_list, int *far_count)
{
    static_assert(KG_NEAR_RINGS == 1 && KG_NT % TEAM == 0 && 32 % TEAM == 0, "team kernel: one near ring, teams inside a warp");
    const GridParams g = *gp;
    const long long i = ((long long)blockIdx.x * KG_NT + threadIdx.x) / TEAM;
    const int sub = threadIdx.x % TEAM;
    const bool live = i < P1;
    float qx = 0.f, qy = 0.f, qz = 0.f;
    if (live) {
        const float x = query[i * 3], y = query[i * 3 + 1], z = query[i * 3 + 2];
        if (T) {
            qx = xadd(xadd(xadd(xmul(T[0], x), xmul(T[1], y)), xmul(T[2], z)), T[3]);
            qy = xadd(xadd(xadd(xmul(T[4], x), xmul(T[5], y)), xmul(T[6], z)), T[7]);
            qz = xadd(xadd(xadd(xmul(T[8], x), xmul(T[9], y)), xmul(T[10], z)), T[11]);
        } else {
            qx = x; qy = y; qz = z;
        }
    }
    float best = INFINITY;
    int bi = 0;
    const bool finite = live && isfinite(qx) && isfinite(qy) && isfinite(qz);
    if (finite) {
        const int cx = cell_coord(qx, g.ox, g.inv_h), cy = cell_coord(qy, g.oy, g.inv_h), cz = cell_coord(qz, g.oz, g.inv_h);
        const int xlo = max(cx - 1, 0), xhi = min(cx + 1, g.nx - 1);
        if (xlo <= xhi) {
            const unsigned xmask = (2u << (xhi - xlo)) - 1u;
#pragma unroll 1
            for (int r = sub; r < 9; r += TEAM) {
                const int z = cz + r / 3 - 1, y = cy + r % 3 - 1;
                if (z < 0 || z >= g.nz || y < 0 || y >= g.ny) continue;
                const int c0 = (z * g.ny + y) * g.nx + xlo;
                unsigned m = __funnelshift_r(bits[c0 >> 5], bits[(c0 >> 5) + 1], c0 & 31) & xmask;
                const int rowid = cell_id(g, 0, y, z);
                while (m) {
                    const int x = xlo + __ffs(m) - 1;
                    const int c = rowid + (x / KG_M) * KG_M3 + x % KG_M;
                    m &= m - 1;
                    for (int j = start[c], e = start[c + 1]; j < e; j++) {
                        const float4 p = sorted[j];
                        const float dx = xsub(qx, p.x), dy2 = xsub(qy, p.y), dz2 = xsub(qz, p.z);
                        const float d2 = xadd(xadd(xmul(dx, dx), xmul(dy2, dy2)), xmul(dz2, dz2));
                        const int pi = __float_as_int(p.w);
                        if (d2 <= best) { if (d2 < best || pi < bi) { best = d2; bi = pi; } }
                    }
                }
            }
        }
    }
    // merge inside the team: smallest distance, lowest index among equals (a lane that saw nothing holds (inf, 0): brute force's answer
    // when no finite distance exists, and never smaller than a found point)
#pragma unroll
    for (int o = TEAM / 2; o > 0; o >>= 1) {
        const float ob = __shfl_xor_sync(0xffffffffu, best, o);
        const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
        if (ob < best || (ob == best && best < INFINITY && oi < bi)) { best = ob; bi = oi; }
    }
    bool decided = !finite;
    if (finite) {
        const float slack = 0.01f + 4e-7f * (float)max(max(g.nx, g.ny), g.nz);
        const float reach = (1.0f - slack) * g.h;
        decided = best <= reach * reach;
    }
    if (live && sub == 0) {
        dist2[i] = best;
        idx[i] = (long long)bi;
    }
    const unsigned und = __ballot_sync(0xffffffffu, !decided && sub == 0);       // warp-aggregated append, one entry per team
    if (und) {
        const int lane = threadIdx.x & 31;
        int base = 0;
        if (lane == 0) base = atomicAdd(far_count, __popc(und));
        base = __shfl_sync(0xffffffffu, base, 0);
        if (!decided && sub == 0) far_list[base + __popc(und & ((1u << lane) - 1))] = (int)i;
    }
}

// Query, phase 2 -- one WARP per far query.  A thread-per-query walk of the rings diverges completely (every lane in its own
// loop nest: measured 17 ms for 240 k queries 9 cm off a 2 M-point surface, ~45 k instructions each).  Here the warp walks
// rings of COARSE cells around its query together: empty coarse cells are skipped, the points of an occupied one -- ONE
// contiguous range of the sorted records, because cell ids are coarse-major -- are dealt out to the lanes as a flat list (round 1
// dealt out the KG_M^3 fine cells, two per lane, each lane in its own point loop: 17 of 32 lanes active, ~50 warp iterations per
// query), and the lanes' results are merged (minimum of (distance, index)) after every coarse cell that was searched.  The search
// stops when the best distance lies inside the fully searched cube, so the result is the brute-force result.
__global__ void __launch_bounds__(KG_NT, KG_FAR_MINB) kg_query_far_kernel(const float *query, const float *T, const GridParams *gp, const int *start,
                                                             const float4 *sorted, const float4 *cbox,
                                                             float *dist2, long long *idx, const int *far_list, const int *far_count)
{
    const GridParams g = *gp;
    const int lane = threadIdx.x & 31;
    const int nfar = *far_count;
    for (int w = blockIdx.x * (KG_NT / 32) + (threadIdx.x >> 5); w < nfar; w += gridDim.x * (KG_NT / 32)) {
        const long long i = far_list[w];
        float qx, qy, qz;
        {
            const float x = query[i * 3], y = query[i * 3 + 1], z = query[i * 3 + 2];
            if (T) {
                qx = xadd(xadd(xadd(xmul(T[0], x), xmul(T[1], y)), xmul(T[2], z)), T[3]);
                qy = xadd(xadd(xadd(xmul(T[4], x), xmul(T[5], y)), xmul(T[6], z)), T[7]);
                qz = xadd(xadd(xadd(xmul(T[8], x), xmul(T[9], y)), xmul(T[10], z)), T[11]);
            } else {
                qx = x; qy = y; qz = z;
            }
        }
        float best = dist2[i];
        int bi = (int)idx[i];
        const int cx = cell_coord(qx, g.ox, g.inv_h), cy = cell_coord(qy, g.oy, g.inv_h), cz = cell_coord(qz, g.oz, g.inv_h);
        const float slack = 0.01f + 4e-7f * (float)max(max(g.nx, g.ny), g.nz);
        const int mx = (g.nx + KG_M - 1) / KG_M, my = (g.ny + KG_M - 1) / KG_M, mz = (g.nz + KG_M - 1) / KG_M;
        auto cdiv = [](int c) { return c >= 0 ? c / KG_M : -((-c + KG_M - 1) / KG_M); };      // floor division
        const int qX = cdiv(cx), qY = cdiv(cy), qZ = cdiv(cz);
        const float hc = g.h * (float)KG_M;
        // rings before the first one that can touch the grid hold nothing
        // (and the walk starts with the 27 cells of box(1) as ONE batch: nearest first works best on a whole neighbourhood)
        int r = max(max(max(-qX, qX - (mx - 1)), max(-qY, qY - (my - 1))), max(max(-qZ, qZ - (mz - 1)), 1));
        int pX0 = 0, pX1 = -1, pY0 = 0, pY1 = -1, pZ0 = 0, pZ1 = -1;      // the part of the grid searched so far (empty)
        for (;; r++) {
            // the coarse cells of ring r inside the grid = box(r) minus box(r - 1), enumerated as up to six slabs (a query far
            // outside the grid would otherwise re-walk everything searched so far in every ring); the lanes test 32 cells for
            // emptiness at a time, then the warp searches the occupied ones one after the other
            const int X0 = max(qX - r, 0), X1 = min(qX + r, mx - 1), Y0 = max(qY - r, 0), Y1 = min(qY + r, my - 1),
                      Z0 = max(qZ - r, 0), Z1 = min(qZ + r, mz - 1);
            const bool fresh = pX1 < pX0 || pY1 < pY0 || pZ1 < pZ0;
            // small rings (box(2) = 125 cells = 4 batches): the whole box, the lanes skipping what box(r - 1) covered -- six
            // thin slabs would be six mostly empty batches
            const bool whole = fresh || r <= 2;
            for (int slab = 0; slab < (whole ? 1 : 6); slab++) {
            int sx0 = X0, sx1 = X1, sy0 = Y0, sy1 = Y1, sz0 = Z0, sz1 = Z1;
            if (!whole) {
                if (slab == 0) sx1 = pX0 - 1;
                if (slab == 1) sx0 = pX1 + 1;
                if (slab >= 2) { sx0 = pX0; sx1 = pX1; }
                if (slab == 2) sy1 = pY0 - 1;
                if (slab == 3) sy0 = pY1 + 1;
                if (slab >= 4) { sy0 = pY0; sy1 = pY1; }
                if (slab == 4) sz1 = pZ0 - 1;
                if (slab == 5) sz0 = pZ1 + 1;
            }
            const int bx = sx1 - sx0 + 1, by = sy1 - sy0 + 1, bz = sz1 - sz0 + 1;
            const int total = (bx > 0 && by > 0 && bz > 0) ? bx * by * bz : 0;
            if (total == 0) continue;
            // t -> (x, y, z) of the slab without integer division (three of them per batch were most of the walk's
            // instructions): quotient = floor((t + 0.5) * rcp(b)).  (t + 0.5) / b is at least 0.5 / b away from an integer and the
            // computed product is within (t + 0.5) / b * 1.3 * 2^-22 of it (reciprocal to 2 ulp, one rounding), so the floor is exact
            // for t < 1.6 M; a slab of more than 2^20 cells (a query far outside a huge flat grid) takes the integer division
            const float ibx = __fdividef(1.0f, (float)bx), iby = __fdividef(1.0f, (float)by);
            const bool exact_rcp = total <= (1 << 20);
            for (int base = 0; base < total; base += 32) {
                const int t = base + lane;
                int rs = 0, re = 0;              // this lane's candidate cell: its range of sorted records
                unsigned key = 0xffffffffu;      // bits of the coarse cell's box distance (>= 0: ordered like the floats); all ones = nothing to search
                if (t < total) {
                    int q1, q2;
                    if (exact_rcp) {
                        q1 = (int)(((float)t + 0.5f) * ibx);
                        q2 = (int)(((float)q1 + 0.5f) * iby);
                    } else {
                        q1 = t / bx;
                        q2 = q1 / by;
                    }
                    const int x = t - q1 * bx, y = q1 - q2 * by;
                    const int X = sx0 + x, Y = sy0 + y, Z = sz0 + q2;
                    const int cc = (Z * my + Y) * mx + X;
                    const bool seen = whole && !fresh && X >= pX0 && X <= pX1 && Y >= pY0 && Y <= pY1 && Z >= pZ0 && Z <= pZ1;
                    const float4 lo = cbox[2 * cc], hi = cbox[2 * cc + 1];      // tight box of the cell's points + their range (kg_coarse_box_kernel)
                    if (!seen && __float_as_int(hi.w) > __float_as_int(lo.w)) {
                        const float ex = fmaxf(fmaxf(lo.x - qx, qx - hi.x), 0.0f);
                        const float ey = fmaxf(fmaxf(lo.y - qy, qy - hi.y), 0.0f);
                        const float ez = fmaxf(fmaxf(lo.z - qz, qz - hi.z), 0.0f);
                        key = __float_as_uint((ex * ex + ey * ey + ez * ez) * 0.9999f);
                        rs = __float_as_int(lo.w);
                        re = __float_as_int(hi.w);
                    }
                }
                // nearest occupied coarse cell of the batch first; `best` is the same in every lane here (merged after each
                // cell), so the batch ends as soon as its nearest unsearched cell is farther than the best distance known --
                // in the first ring that touches the surface this drops most of the up to 32 candidates unsearched
                for (;;) {
                    const unsigned mk = __reduce_min_sync(0xffffffffu, key);
                    if (mk == 0xffffffffu || __uint_as_float(mk) > best) break;
                    const int srcl = __ffs(__ballot_sync(0xffffffffu, key == mk)) - 1;
                    if (lane == srcl) key = 0xffffffffu;
                    // the cell's points are one contiguous range of the sorted records (coarse-major cell ids): a flat list, 32
                    // points per step, two steps in flight -- every lane works, nothing is looked up per fine cell
                    const int s0 = __shfl_sync(0xffffffffu, rs, srcl), e0 = __shfl_sync(0xffffffffu, re, srcl);
                    for (int j = s0 + lane; j < e0; j += 64) {
                        const bool two = j + 32 < e0;
                        const float4 p = sorted[j];
                        const float4 p2 = sorted[two ? j + 32 : j];
                        {
                            const float dx = xsub(qx, p.x), dy = xsub(qy, p.y), dz = xsub(qz, p.z);
                            const float d2 = xadd(xadd(xmul(dx, dx), xmul(dy, dy)), xmul(dz, dz));
                            const int pi = __float_as_int(p.w);
                            if (d2 <= best) { if (d2 < best || pi < bi) { best = d2; bi = pi; } }
                        }
                        {
                            const float dx = xsub(qx, p2.x), dy = xsub(qy, p2.y), dz = xsub(qz, p2.z);
                            const float d2 = xadd(xadd(xmul(dx, dx), xmul(dy, dy)), xmul(dz, dz));
                            const int pi = __float_as_int(p2.w);
                            if (d2 <= best) { if (d2 < best || pi < bi) { best = d2; bi = pi; } }
                        }
                    }
                    // merge: minimum of (distance, index) over the lanes, known to all of them (distances are >= 0 or +inf:
                    // their bit patterns order like the values; indices are non-negative)
                    const unsigned mb = __reduce_min_sync(0xffffffffu, __float_as_uint(best));
                    bi = (int)__reduce_min_sync(0xffffffffu, __float_as_uint(best) == mb ? (unsigned)bi : 0xffffffffu);
                    best = __uint_as_float(mb);
                }
            }
            }      // slabs
            pX0 = X0; pX1 = X1; pY0 = Y0; pY1 = Y1; pZ0 = Z0; pZ1 = Z1;
            const float reach = ((float)r - slack) * hc;
            if (reach > 0.0f && best <= reach * reach) break;
            if (qX - r <= 0 && qX + r >= mx - 1 && qY - r <= 0 && qY + r >= my - 1 && qZ - r <= 0 && qZ + r >= mz - 1) break;
        }
        if (lane == 0) {
            dist2[i] = best;
            idx[i] = (long long)bi;
        }
    }
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

// coarse cells: at most ceil(n / KG_M) per axis; n_x n_y n_z <= KG_MAX_CELLS bounds the product by KG_MAX_CELLS / KG_M^3 plus edges
constexpr size_t KG_COARSE = (size_t)KG_MAX_CELLS / (KG_M * KG_M * KG_M) + 3 * ((size_t)KG_MAX_CELLS / (KG_M * KG_M)) + 4096;
constexpr size_t KG_PADDED = ((size_t)KG_MAX_CELLS / 4096 + 1) * 4096;      // histogram length: cells + 1, padded to whole chunks

size_t e2e_knn1_grid_workspace_bytes(long long P2)
{
    if (P2 < 0) P2 = 0;
    return 2 * kg_a256(KG_PADDED * 4) + kg_a256((size_t)P2 * 4) + kg_a256((size_t)P2 * 16) + kg_a256(sizeof(GridParams)) +
           kg_a256((KG_PADDED / 4096) * 4) + kg_a256(KG_COARSE * 4) + kg_a256(KG_PADDED / 8) + 256 + 256 + 256 + kg_a256(KG_COARSE * 32);
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
    int *start = (int *)w;              w += kg_a256(KG_PADDED * 4);
    int *cursor = (int *)w;             w += kg_a256(KG_PADDED * 4);
    int *cell_of = (int *)w;            w += kg_a256((size_t)P2 * 4);
    float4 *sorted = (float4 *)w;       w += kg_a256((size_t)P2 * 16);
    GridParams *gp = (GridParams *)w;   w += kg_a256(sizeof(GridParams));
    int *chunk_sum = (int *)w;          w += kg_a256((KG_PADDED / 4096) * 4);
    int *coarse = (int *)w;             w += kg_a256(KG_COARSE * 4);
    unsigned *bits = (unsigned *)w;     w += kg_a256(KG_PADDED / 8);
    unsigned *bb = (unsigned *)w;       w += 3 * 256;      // bbox words, far-query counter
    float4 *cbox = (float4 *)w;         // box + record range of every coarse cell (kg_coarse_box_kernel writes all of them: no clear)
    // bbox accumulators: minima start at all ones, maxima at zero (ordered encoding); the histogram's padding stays zero
    if (cudaMemsetAsync(bb, 0xff, 12, st) != cudaSuccess || cudaMemsetAsync(bb + 3, 0x00, 12, st) != cudaSuccess ||
        cudaMemsetAsync(start, 0, KG_PADDED * 4, st) != cudaSuccess || cudaMemsetAsync(coarse, 0, KG_COARSE * 4 , st) != cudaSuccess ||
        cudaMemsetAsync(bits, 0, KG_PADDED / 8, st) != cudaSuccess)
        return finish_launch("knn1_grid: memset");
    const int nb = kg_blocks(P2), chunks = (int)(KG_PADDED / 4096);
    kg_bbox_kernel<<<nb < kNumSMs * 8 ? nb : kNumSMs * 8, KG_NT, 0, st>>>(ref, P2, bb);
    kg_params_kernel<<<1, 1, 0, st>>>(bb, P2, gp);
    for (int pass = 0; pass < KG_PASSES; pass++) {
        if (pass) kg_clear_kernel<<<kNumSMs * 8, KG_NT, 0, st>>>((int4 *)start, (int)(KG_PADDED / 4), gp);
        kg_count_kernel<<<nb, KG_NT, 0, st>>>(ref, P2, gp, cell_of, start);
        kg_refine_kernel<<<1, 1, 0, st>>>(gp, P2, pass == KG_PASSES - 1);
    }
    kg_scan1_kernel<<<chunks, 1024, 0, st>>>(start, chunk_sum, gp);
    kg_scan2_kernel<<<1, 1024, 0, st>>>(chunk_sum, gp);
    kg_scan3_kernel<<<chunks, 1024, 0, st>>>(start, chunk_sum, cursor, gp);
    kg_fill_kernel<<<nb, KG_NT, 0, st>>>(ref, P2, cell_of, cursor, sorted);
    kg_coarse_kernel<<<nb, KG_NT, 0, st>>>(P2, gp, cell_of, coarse, bits);
    kg_coarse_box_kernel<<<kNumSMs * 16, KG_NT, 0, st>>>(gp, start, sorted, coarse, cbox);
    count_launch(2 + 3 * KG_PASSES - 1 + 6);
    return finish_launch("knn1_grid_build");
}

int e2e_knn1_grid_query(const float *query, const float *transform, long long P1, long long P2,
                        float *dist2, long long *idx, const void *workspace, void *stream)
{
    cudaStream_t st = (cudaStream_t)stream;
    E2E_REQUIRE(query && dist2 && idx && workspace && P1 > 0 && P2 > 0, "knn1_grid_query: bad arguments");
    const unsigned char *w = (const unsigned char *)workspace;
    const int *start = (const int *)w;          w += 2 * kg_a256(KG_PADDED * 4) + kg_a256((size_t)P2 * 4);
    const float4 *sorted = (const float4 *)w;   w += kg_a256((size_t)P2 * 16);
    const GridParams *gp = (const GridParams *)w;   w += kg_a256(sizeof(GridParams)) + kg_a256((KG_PADDED / 4096) * 4);
    const int *coarse = (const int *)w;             w += kg_a256(KG_COARSE * 4);
    const unsigned *bits = (const unsigned *)w;
    int *far_count = (int *)(const_cast<unsigned char *>(w) + kg_a256(KG_PADDED / 8) + 256);      // after the bitmap and the bbox words
    const float4 *cbox = (const float4 *)(w + kg_a256(KG_PADDED / 8) + 3 * 256);
    const long long qb = (P1 + KG_NT - 1) / KG_NT;
    E2E_REQUIRE(qb < (1ll << 31) && P1 < (1ll << 31), "knn1_grid: too many query points");
    // the list of far queries borrows the first P1 ints of the cursor array (dead after the build: 2^24 ints)
    E2E_REQUIRE(P1 <= (long long)KG_PADDED, "knn1_grid: more than %zu query points per call", KG_PADDED);
    int *far_list = (int *)((unsigned char *)const_cast<void *>(workspace) + kg_a256(KG_PADDED * 4));
    if (cudaMemsetAsync(far_count, 0, 4, st) != cudaSuccess) return finish_launch("knn1_grid_query: memset");
    // few queries (ICP): a team of 8 lanes per query; E2E_KNN_TEAM=0 / 1 forces the one-thread / team kernel
    constexpr int TEAM = 8;
    static const int team_env = [] { const char *e = getenv("E2E_KNN_TEAM"); return e ? atoi(e) : -1; }();
    const bool team = team_env < 0 ? P1 <= (1ll << 16) : team_env != 0;
    if (team) {
        const long long tb = (P1 * TEAM + KG_NT - 1) / KG_NT;
        E2E_REQUIRE(tb < (1ll << 31), "knn1_grid: too many query points");
        kg_query_near_team_kernel<TEAM><<<(unsigned)tb, KG_NT, 0, st>>>(query, transform, P1, gp, start, sorted, bits, dist2, idx, far_list, far_count);
    } else {
        kg_query_near_kernel<<<(unsigned)qb, KG_NT, 0, st>>>(query, transform, P1, gp, start, sorted, bits, dist2, idx, far_list, far_count);
    }
    long long fb = (P1 + KG_NT / 32 - 1) / (KG_NT / 32);
    if (fb > kNumSMs * 32) fb = kNumSMs * 32;
    kg_query_far_kernel<<<(unsigned)fb, KG_NT, 0, st>>>(query, transform, gp, start, sorted, cbox, dist2, idx, far_list, far_count);
    count_launch(2);
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
