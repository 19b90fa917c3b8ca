// Integrator hot path (modules/pipeline.py:137-171 + modules/integrator.py:15-126) on sm_100a.
//
// The reference sums, per voxel, the contributions w and w*v of every (ray,sample,corner)
// entry that lands in it, in ascending entry order e (CPU index_add_ order),
// in fp32.  fp32 addition is not associative and the result is stored as fp16, so bit-exact
// parity needs exactly that order.  The work is split in two:
//
// PLAN (needs only the geometry: per-ray records + depth mask -- independent of the network, so the pipeline builds
// it on a side stream while AdapNet++ / FusionNet run):
//   count    : one thread per (ray,sample): recompute the 8 corners from the extractor's per-ray
//              record, find/claim the voxel's slot in an open-addressing hash (key = linear voxel
//              index) and take an arrival number from its counter; slot and arrival number are
//              remembered per entry (coalesced 64 B/thread).  The thread that claims a voxel
//              appends the slot to its block's touched list.
//   offsets  : one thread per touched voxel: block scan of the counts + one atomic per block gives
//              every voxel a contiguous segment {off, len}; voxels with more than 32 entries are queued.
//   scatter  : one thread per (ray,sample): write e to arrival[off + arrival number].
//   rank     : one thread per (ray,sample): for each of its entries count the entries of the same voxel with a smaller e
//              and store {e, w} at sorted[off + rank]; over-long voxels (> 2048 entries) are left for a block sort.
// APPLY (the only part after the network; one kernel):
//   apply    : one thread per touched voxel streams its sorted segment; the network value / label of every entry is
//              gathered by e: fp32 sums in ascending e, the running-mean update with fp16 round-to-nearest stores
//              (integrator.py:77-88), the semantic "highest entry wins" update (integrator.py:90-124, App. A.5), and
//              the slot goes back to idle.  Over-long voxels: a block sorts the segment first.
//
// No floating-point atomics, no G^3 scratch, no global sort.  The result is deterministic and bit-identical to the
// single-threaded reference for every list length (near-camera frames put >10^4 entries into one voxel; see tests).
#include "ojdf_internal.h"

namespace ojdf {

constexpr int kThreads = 256;
constexpr int kSegment = kThreads * 8;        // entries per block == max voxels a block can claim
constexpr int kRankMax = 2048;                // longest per-voxel list ranked by counting; above: block sort
constexpr uint32_t kNoSlot = 0xFFFFFFFFu;
constexpr int kCoopBlocks = 148;              // blocks draining the long-voxel queue

struct Workspace {
    uint4 *table;          // {key + 1 (0 = empty), arrival counter, segment offset, segment length}; zero when idle
    uint32_t *eslot;       // slot of every entry (kNoSlot = out of grid / masked)
    uint32_t *earr;        // arrival number of every entry inside its voxel
    uint32_t *list;        // touched slots, per-block segments of kSegment
    uint32_t *arrival;     // e of every entry, grouped by voxel, arrival order
    uint2 *sorted;         // {e, weight bits} records, grouped by voxel, ascending e
    uint32_t *queue;       // slots of voxels with more than kRankMax entries
    uint32_t *count;       // touched voxels per block
    uint32_t *ctrl;        // [0] segment bump cursor, [1] queue length; zero when idle
    uint32_t slots_mask;
    int log2_slots;
};

static inline int log2_slots_for(long long cap)
{
    long long want = cap + cap / 2 + 1024;        // load factor <= 2/3 even if every entry is a different voxel
    int l = 10;
    while ((1ll << l) < want) ++l;
    return l;
}

static size_t align256(size_t x) { return (x + 255) & ~(size_t)255; }

// Lay the workspace out for `cap` entries; returns the bytes needed (base may be null: sizing only).
static size_t carve(Workspace &ws, void *base, long long cap)
{
    const int l = log2_slots_for(cap);
    const long long blocks = cap / kSegment + 1;
    const uintptr_t b = (uintptr_t)base;
    size_t off = 0;
    auto take = [&](size_t bytes) { uintptr_t p = b + off; off += align256(bytes); return p; };
    ws.ctrl = (uint32_t *)take(256);
    ws.table = (uint4 *)take(sizeof(uint4) << l);          // [0, here) must be zero between calls
    ws.eslot = (uint32_t *)take(4 * (size_t)cap);
    ws.earr = (uint32_t *)take(4 * (size_t)cap);
    ws.list = (uint32_t *)take(4 * (size_t)blocks * kSegment);
    ws.arrival = (uint32_t *)take(4 * (size_t)cap);
    ws.sorted = (uint2 *)take(8 * (size_t)cap);
    ws.queue = (uint32_t *)take(4 * ((size_t)cap / kRankMax + 1));
    ws.count = (uint32_t *)take(4 * (size_t)blocks);
    ws.log2_slots = l;
    ws.slots_mask = (uint32_t)((1ull << l) - 1);
    return off;
}

static size_t idle_prefix_bytes(long long cap) { return align256(256) + align256(sizeof(uint4) << log2_slots_for(cap)); }

static long long round_up_entries(long long e) { return (e + kSegment - 1) / kSegment * kSegment; }

// The layout is a function of the workspace SIZE only, so calls with different frame sizes on one
// workspace all see the same (idle) hash table.
static long long capacity_for_bytes(size_t bytes)
{
    long long lo = 0, hi = (1ll << 31) / kSegment - 1;
    Workspace tmp;
    while (lo < hi) {
        const long long mid = (lo + hi + 1) / 2;
        if (carve(tmp, nullptr, mid * kSegment) <= bytes) lo = mid; else hi = mid - 1;
    }
    return lo * kSegment;
}

static int bind_workspace(Workspace &ws, void *base, size_t bytes, long long entries)
{
    const long long cap = capacity_for_bytes(bytes);
    if (cap < round_up_entries(entries) || cap == 0) return OJDF_ERR_WORKSPACE;
    carve(ws, base, cap);
    return 0;
}

// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t hash_slot(uint32_t key, int log2_slots)
{
    return (key * 2654435761u) >> (32 - log2_slots);
}

// The 8 (voxel key, weight) pairs of one (ray,sample); key = kNoSlot marks an out-of-grid corner.
struct Sample {
    uint32_t key[8];
    float w[8];
};

__device__ __forceinline__ void sample_corners(const double *__restrict__ ray, long long n, int i, int X, int Y, int Z, Sample &s)
{
    const double2 *rp = reinterpret_cast<const double2 *>(ray + 6 * n);
    const double2 r0 = __ldg(rp), r1 = __ldg(rp + 1), r2 = __ldg(rp + 2);
    const Axis ax = axis_setup(ray_sample(r0.x, r1.y, i));
    const Axis ay = axis_setup(ray_sample(r0.y, r2.x, i));
    const Axis az = axis_setup(ray_sample(r1.x, r2.y, i));
#pragma unroll
    for (int c = 0; c < 8; ++c) {
        long long ix, iy, iz;
        const bool ok = corner_index(ax, ay, az, c, X, Y, Z, ix, iy, iz);
        s.key[c] = ok ? (uint32_t)((ix * Y + iy) * (long long)Z + iz) : kNoSlot;
        s.w[c] = (float)corner_weight(ax, ay, az, c);                   // .float(), integrator.py:45
    }
}

__device__ __forceinline__ void sample_updates(const long long *__restrict__ idx, const double *__restrict__ wts, long long m,
                                               int X, int Y, int Z, Sample &s)
{
#pragma unroll
    for (int c = 0; c < 8; ++c) {
        const long long *ip = idx + (m * 8 + c) * 3;
        const long long ix = ip[0], iy = ip[1], iz = ip[2];
        const bool ok = ix >= 0 && ix < X && iy >= 0 && iy < Y && iz >= 0 && iz < Z;
        s.key[c] = ok ? (uint32_t)((ix * Y + iy) * (long long)Z + iz) : kNoSlot;
        s.w[c] = (float)wts[m * 8 + c];
    }
}

// Source of the entries: FRAME = recompute from the per-ray record (T samples per ray, masked rays
// skipped), otherwise the reference's materialised `updates` tensors (one thread per sample).
struct Source {
    const double *ray; const float *filt; const float *est;      // frame form
    const float *values; const long long *idx; const double *wts; // updates form
    long long items;                                             // N*T or M1
    int P, T; float clampv; int X, Y, Z;
};

template <bool FRAME>
__device__ __forceinline__ bool load_sample(const Source &src, long long t, Sample &s, float &val)
{
    if (t >= src.items) return false;
    if (FRAME) {
        const long long n = t / src.T;
        const int k = (int)(t - n * src.T);
        if (!(src.filt[n] != 0.0f)) return false;                       // modules/pipeline.py:143-146
        sample_corners(src.ray, n, k - src.P / 2, src.X, src.Y, src.Z, s);
        val = 0.0f;                                                      // the plan kernels never look at the values
    } else {
        sample_updates(src.idx, src.wts, t, src.X, src.Y, src.Z, s);
        val = 0.0f;
    }
    return true;
}

__device__ __forceinline__ void store8(uint32_t *dst, const uint32_t v[8])
{
    uint4 *o = reinterpret_cast<uint4 *>(dst);
    o[0] = make_uint4(v[0], v[1], v[2], v[3]);
    o[1] = make_uint4(v[4], v[5], v[6], v[7]);
}

__device__ __forceinline__ void load8(const uint32_t *src, uint32_t v[8])
{
    const uint4 *i = reinterpret_cast<const uint4 *>(src);
    const uint4 a = i[0], b = i[1];
    v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
}

template <bool FRAME>
__global__ void __launch_bounds__(kThreads)
count_kernel(Source src, Workspace ws)
{
    __shared__ uint32_t s_count;
    if (threadIdx.x == 0) s_count = 0;
    __syncthreads();
    uint32_t *seg_list = ws.list + (size_t)blockIdx.x * kSegment;
    const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (t < src.items) {
        Sample s;
        float val;
        uint32_t slot[8], arr[8];
#pragma unroll
        for (int c = 0; c < 8; ++c) { slot[c] = kNoSlot; arr[c] = 0; }
        if (load_sample<FRAME>(src, t, s, val)) {
            uint32_t cur[8];
#pragma unroll
            for (int c = 0; c < 8; ++c) {                               // 8 independent first probes in flight
                slot[c] = hash_slot(s.key[c], ws.log2_slots);
                cur[c] = s.key[c] != kNoSlot ? __ldcg(&ws.table[slot[c]].x) : 0u;
            }
            bool is_new[8];
#pragma unroll
            for (int c = 0; c < 8; ++c) {
                is_new[c] = false;
                if (s.key[c] == kNoSlot) { slot[c] = kNoSlot; continue; }
                const uint32_t k1 = s.key[c] + 1u;
                uint32_t sl = slot[c], cu = cur[c];
                while (true) {                                          // linear probing; CAS only on empty slots
                    if (cu == 0u) {
                        cu = atomicCAS(&ws.table[sl].x, 0u, k1);
                        if (cu == 0u) { is_new[c] = true; break; }
                    }
                    if (cu == k1) break;
                    sl = (sl + 1) & ws.slots_mask;
                    cu = __ldcg(&ws.table[sl].x);
                }
                slot[c] = sl;
            }
#pragma unroll
            for (int c = 0; c < 8; ++c)                                 // 8 independent counter bumps in flight
                if (slot[c] != kNoSlot) arr[c] = atomicAdd(&ws.table[slot[c]].y, 1u);
#pragma unroll
            for (int c = 0; c < 8; ++c)
                if (is_new[c]) seg_list[atomicAdd(&s_count, 1u)] = slot[c];
        }
        store8(ws.eslot + t * 8, slot);
        store8(ws.earr + t * 8, arr);
    }
    __syncthreads();
    if (threadIdx.x == 0) ws.count[blockIdx.x] = s_count;
}

// Block-wide exclusive scan of one uint per thread (kThreads threads); returns the block total.
__device__ __forceinline__ uint32_t block_exclusive_scan(uint32_t v, uint32_t &total, uint32_t *s_warp /* [8] */)
{
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    uint32_t inc = v;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const uint32_t o = __shfl_up_sync(0xffffffffu, inc, d);
        if (lane >= d) inc += o;
    }
    if (lane == 31) s_warp[warp] = inc;
    __syncthreads();
    uint32_t base = 0, tot = 0;
#pragma unroll
    for (int i = 0; i < kThreads / 32; ++i) { if (i < warp) base += s_warp[i]; tot += s_warp[i]; }
    __syncthreads();
    total = tot;
    return base + inc - v;
}

__global__ void __launch_bounds__(kThreads)
offsets_kernel(Workspace ws)
{
    __shared__ uint32_t s_warp[kThreads / 32];
    __shared__ uint32_t s_base;
    const uint32_t cnt = ws.count[blockIdx.x];
    const size_t seg0 = (size_t)blockIdx.x * kSegment;
    constexpr int R = kSegment / kThreads;
    uint32_t len[R], slot[R], mine = 0;
#pragma unroll
    for (int r = 0; r < R; ++r) {                                       // thread owns R consecutive list positions
        const uint32_t j = threadIdx.x * R + r;
        len[r] = 0; slot[r] = 0;
        if (j < cnt) { slot[r] = ws.list[seg0 + j]; len[r] = __ldcg(&ws.table[slot[r]].y); }
        mine += len[r];
    }
    uint32_t total;
    uint32_t pre = block_exclusive_scan(mine, total, s_warp);
    if (threadIdx.x == 0) s_base = total ? atomicAdd(&ws.ctrl[0], total) : 0u;
    __syncthreads();
    pre += s_base;
    const uint32_t lane = threadIdx.x & 31, lt_mask = (1u << lane) - 1u;
#pragma unroll
    for (int r = 0; r < R; ++r) {
        const uint32_t j = threadIdx.x * R + r;
        const bool valid = j < cnt;
        if (valid) {
            uint2 *zw = reinterpret_cast<uint2 *>(&ws.table[slot[r]].z);
            *zw = make_uint2(pre, len[r]);                              // segment {off, len}
            pre += len[r];
        }
        // queue the over-long voxels (warp-aggregated: one atomic per warp)
        const bool q = valid && len[r] > (uint32_t)kRankMax;
        const uint32_t m = __ballot_sync(0xffffffffu, q);
        if (m) {
            const int leader = __ffs(m) - 1;
            uint32_t base = 0;
            if ((int)lane == leader) base = atomicAdd(&ws.ctrl[1], (uint32_t)__popc(m));
            base = __shfl_sync(0xffffffffu, base, leader);
            if (q) ws.queue[base + __popc(m & lt_mask)] = slot[r];
        }
    }
}

// arrival[off + arrival number] = e
__global__ void __launch_bounds__(kThreads)
scatter_kernel(long long items, Workspace ws)
{
    const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= items) return;
    uint32_t slot[8], arr[8], off[8];
    load8(ws.eslot + t * 8, slot);
    load8(ws.earr + t * 8, arr);
#pragma unroll
    for (int c = 0; c < 8; ++c) off[c] = slot[c] != kNoSlot ? __ldcg(&ws.table[slot[c]].z) : 0u;
#pragma unroll
    for (int c = 0; c < 8; ++c)
        if (slot[c] != kNoSlot) ws.arrival[off[c] + arr[c]] = (uint32_t)(t * 8 + c);
}

// One thread per (ray,sample): for each of its entries count the entries of the same voxel with a smaller e (a dense,
// independent, cache-friendly loop over the voxel's arrival segment) and store {e, weight} at sorted[off + rank].
// Entries of over-long voxels keep their arrival position; a whole block sorts that segment later.  Needs only the
// geometry, so it belongs to the plan.
template <bool FRAME>
__global__ void __launch_bounds__(kThreads)
rank_kernel(Source src, Workspace ws)
{
    const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    Sample s;
    float val;
    if (!load_sample<FRAME>(src, t, s, val)) return;
    uint32_t slot[8], arr[8];
    load8(ws.eslot + t * 8, slot);
    load8(ws.earr + t * 8, arr);
    uint2 seg[8];
#pragma unroll
    for (int c = 0; c < 8; ++c)
        seg[c] = slot[c] != kNoSlot ? __ldcg(reinterpret_cast<const uint2 *>(&ws.table[slot[c]].z)) : make_uint2(0u, 0u);
#pragma unroll
    for (int c = 0; c < 8; ++c) {
        if (slot[c] == kNoSlot) continue;
        const uint32_t e = (uint32_t)(t * 8 + c), off = seg[c].x, len = seg[c].y;
        uint32_t r = arr[c];                                            // over-long voxel: keep arrival order, sorted later
        if (len <= (uint32_t)kRankMax) {
            const uint32_t *a = ws.arrival + off;
            uint32_t r0 = 0, r1 = 0, r2 = 0, r3 = 0, q = 0;
            for (; q + 4 <= len; q += 4) {                              // rank = number of smaller entries of this voxel
                r0 += a[q] < e; r1 += a[q + 1] < e; r2 += a[q + 2] < e; r3 += a[q + 3] < e;
            }
            for (; q < len; ++q) r0 += a[q] < e;
            r = (r0 + r1) + (r2 + r3);
        }
        ws.sorted[off + r] = make_uint2(e, __float_as_uint(s.w[c]));
    }
}

// ---------------------------------------------------------------------------------------------
// The second half: everything that needs the network output.  `est` / labels / scores are gathered per entry from
// the entry number e (frame form: ray n = e / (8 T), sample k = (e / 8) % T; updates form: record e / 8).
struct Apply {
    __half *tsdf; __half *wvol; uint8_t *ids; __half *scores;
    const float *est;              // frame form: (N, P) network output; updates form: (M1) values
    const uint8_t *sem_ids; const float *sem_scores;
    uint32_t T8;                   // frame form: 8 * tail; updates form: 8
    int P, frame, do_sem;
    float clampv;
};

__device__ __forceinline__ float entry_value(const Apply &a, uint32_t e)
{
    if (!a.frame) return a.est[e >> 3];
    const uint32_t n = e / a.T8, k = (e - n * a.T8) >> 3;
    const float v = a.est[(size_t)n * a.P + k], c = a.clampv;
    return v < -c ? -c : (v > c ? c : v);                                  // torch.clamp, pipeline.py:157-159
}
__device__ __forceinline__ uint32_t entry_record(const Apply &a, uint32_t e) { return a.frame ? e / a.T8 : e >> 3; }

// Running state of one voxel while its entries are visited in ascending e (integrator.py:55-124, App. A.4-A.5).
struct VoxelAcc {
    float W, U;
    uint32_t e_last, e_label;
    __device__ __forceinline__ void init() { W = 0.0f; U = 0.0f; e_last = 0; e_label = kNoSlot; }
    __device__ __forceinline__ void add(float w, float u, uint32_t e, bool label_differs)
    {
        W = __fadd_rn(W, w);                                                // index_add_, ascending e (integrator.py:60,65)
        U = __fadd_rn(U, u);
        e_last = e;
        if (label_differs) e_label = e;
    }
};

__device__ __forceinline__ void store_voxel(const Apply &a, uint32_t key, const VoxelAcc &v, float wo, float vo, uint8_t id_old, float sc_old)
{
    const float wn = __fadd_rn(wo, v.W);
    a.wvol[key] = __float2half_rn(wn);                                                        // integrator.py:77-78
    a.tsdf[key] = __float2half_rn(__fdiv_rn(__fadd_rn(__fmul_rn(wo, vo), v.U), wn));          // integrator.py:82-83 (0/0 -> NaN kept)
    if (a.do_sem) {
        const float s_last = a.sem_scores[entry_record(a, v.e_last)];                         // highest entry wins
        a.scores[key] = __float2half_rn(s_last > sc_old ? s_last : sc_old);                   // integrator.py:112-113,124
        if (v.e_label != kNoSlot) {                                                           // some entry's label differs
            const uint32_t r = entry_record(a, v.e_label);
            a.ids[key] = a.sem_scores[r] > sc_old ? a.sem_ids[r] : id_old;                     // integrator.py:115-116,123
        }
    }
}

// All-ascending bitonic network over seg[0..len): comparators whose upper index is >= len are
// no-ops (virtual +inf padding), so any length sorts in place.
__device__ __forceinline__ void block_sort(uint2 *seg, uint32_t len)
{
    uint32_t P = 1;
    while (P < len) P <<= 1;
    for (uint32_t k = 2; k <= P; k <<= 1) {
        for (uint32_t i = threadIdx.x; i < P; i += blockDim.x) {        // mirror step: i <-> i ^ (k-1)
            const uint32_t j = i ^ (k - 1);
            if (j > i && j < len) {
                const uint2 x = seg[i], y = seg[j];
                if (x.x > y.x) { seg[i] = y; seg[j] = x; }
            }
        }
        __syncthreads();
        for (uint32_t s = k >> 2; s >= 1; s >>= 1) {                    // half-cleaners: i <-> i ^ s
            for (uint32_t i = threadIdx.x; i < P; i += blockDim.x) {
                const uint32_t j = i ^ s;
                if (j > i && j < len) {
                    const uint2 x = seg[i], y = seg[j];
                    if (x.x > y.x) { seg[i] = y; seg[j] = x; }
                }
            }
            __syncthreads();
        }
    }
}

// One thread: visit the sorted segment of the voxel in `slot` (4 records and their gathers in flight at a time),
// update the volumes, free the slot.
__device__ __forceinline__ void apply_voxel(const Workspace &ws, const Apply &a, uint32_t slot, const uint4 entry)
{
    const uint32_t key = entry.x - 1u, off = entry.z, len = entry.w;
    ws.table[slot] = make_uint4(0u, 0u, 0u, 0u);                        // slot back to idle for the next frame
    uint8_t id_old = 0;
    float sc_old = 0.0f;
    if (a.do_sem) { id_old = a.ids[key]; sc_old = __half2float(a.scores[key]); }
    const float wo = __half2float(a.wvol[key]), vo = __half2float(a.tsdf[key]);
    VoxelAcc acc;
    acc.init();
    const uint2 *seg = ws.sorted + off;
    uint32_t q = 0;
    for (; q + 4 <= len; q += 4) {
        uint2 r[4];
        float u[4];
        uint32_t lab[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) r[i] = seg[q + i];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            u[i] = __fmul_rn(__uint_as_float(r[i].y), entry_value(a, r[i].x));      // integrator.py:55, separately rounded
            lab[i] = a.do_sem ? (uint32_t)a.sem_ids[entry_record(a, r[i].x)] : 0u;
        }
#pragma unroll
        for (int i = 0; i < 4; ++i) acc.add(__uint_as_float(r[i].y), u[i], r[i].x, a.do_sem && lab[i] != (uint32_t)id_old);
    }
    for (; q < len; ++q) {
        const uint2 r = seg[q];
        const float w = __uint_as_float(r.y);
        const bool differs = a.do_sem && (uint32_t)a.sem_ids[entry_record(a, r.x)] != (uint32_t)id_old;
        acc.add(w, __fmul_rn(w, entry_value(a, r.x)), r.x, differs);
    }
    store_voxel(a, key, acc, wo, vo, id_old, sc_old);
}

// grid = kCoopBlocks blocks that sort + apply the over-long voxels (scheduled first: they are the long poles) followed
// by one block per count block (one thread per touched voxel).
__global__ void __launch_bounds__(kThreads)
apply_kernel(Workspace ws, Apply a)
{
    if (blockIdx.x < (uint32_t)kCoopBlocks) {
        const uint32_t nq = ws.ctrl[1];
        for (uint32_t q = blockIdx.x; q < nq; q += kCoopBlocks) {
            const uint32_t slot = ws.queue[q];
            const uint4 entry = ws.table[slot];
            block_sort(ws.sorted + entry.z, entry.w);
            if (threadIdx.x == 0) apply_voxel(ws, a, slot, entry);
            __syncthreads();
        }
        return;
    }
    const uint32_t sb = blockIdx.x - kCoopBlocks;
    const uint32_t cnt = ws.count[sb];
    const size_t seg0 = (size_t)sb * kSegment;
    for (uint32_t j = threadIdx.x; j < cnt; j += blockDim.x) {
        const uint32_t slot = ws.list[seg0 + j];
        const uint4 entry = __ldcg(&ws.table[slot]);
        // over-long voxels belong to the cooperative blocks, which may already have applied the voxel and put its slot
        // back to idle ({0,0,0,0}) by the time this block runs: never act on an idle slot
        if (entry.x == 0u || entry.w > (uint32_t)kRankMax) continue;
        apply_voxel(ws, a, slot, entry);
    }
}

// Runs after the apply kernel (stream order): control words back to idle.
__global__ void reset_ctrl_kernel(Workspace ws) { if (threadIdx.x < 4) ws.ctrl[threadIdx.x] = 0u; }

// Plan: everything that only needs the geometry (per-ray records + mask, or the updates' indices / weights).
template <bool FRAME>
static int run_plan(const Source &src, Workspace &ws, cudaStream_t s)
{
    const unsigned blocks = (unsigned)((src.items + kThreads - 1) / kThreads);
    count_kernel<FRAME><<<blocks, kThreads, 0, s>>>(src, ws);
    offsets_kernel<<<blocks, kThreads, 0, s>>>(ws);
    scatter_kernel<<<blocks, kThreads, 0, s>>>(src.items, ws);
    rank_kernel<FRAME><<<blocks, kThreads, 0, s>>>(src, ws);
    return launched(4);
}

static int run_apply(long long items, const Apply &a, Workspace &ws, cudaStream_t s)
{
    const unsigned blocks = (unsigned)((items + kThreads - 1) / kThreads);
    apply_kernel<<<blocks + kCoopBlocks, kThreads, 0, s>>>(ws, a);
    reset_ctrl_kernel<<<1, 32, 0, s>>>(ws);
    return launched(2);
}

}  // namespace ojdf

using namespace ojdf;

extern "C" size_t ojdf_integrate_workspace_bytes(int64_t max_entries)
{
    if (max_entries <= 0 || max_entries >= 0x7FFFFFFFll - kSegment) return 0;
    Workspace ws;
    return carve(ws, nullptr, round_up_entries(max_entries));
}

extern "C" size_t ojdf_integrate_workspace_idle_bytes(size_t workspace_bytes)
{
    const long long cap = capacity_for_bytes(workspace_bytes);
    return cap ? idle_prefix_bytes(cap) : 0;
}

extern "C" int ojdf_integrate_workspace_init(void *workspace_dev, size_t workspace_bytes, void *stream)
{
    if (!workspace_dev || workspace_bytes == 0) return OJDF_ERR_WORKSPACE;
    const long long cap = capacity_for_bytes(workspace_bytes);
    if (cap == 0) return OJDF_ERR_WORKSPACE;
    const cudaError_t e = cudaMemsetAsync(workspace_dev, 0, idle_prefix_bytes(cap), (cudaStream_t)stream);
    return e == cudaSuccess ? 0 : (int)e;
}

static int check_common(const void *tsdf, const void *wvol, int X, int Y, int Z, const uint8_t *ids, const float *scores,
                        const uint8_t *ids_vol, const void *scores_vol, int do_sem, const void *ws, long long entries)
{
    if (!tsdf || !wvol || X <= 0 || Y <= 0 || Z <= 0) return OJDF_ERR_BADARG;
    if (do_sem && (!ids || !scores || !ids_vol || !scores_vol)) return OJDF_ERR_BADARG;
    if (!ws) return OJDF_ERR_WORKSPACE;
    if ((long long)X * Y * Z >= 0xFFFFFFFFll || entries >= 0x7FFFFFFFll - kSegment) return OJDF_ERR_TOOLARGE;
    return 0;
}

static int frame_args_ok(int64_t N, int P, int tail)
{
    return !(N < 0 || P < 1 || P > 33 || !(P & 1) || tail < 1 || tail > P);
}

extern "C" int ojdf_integrate_plan(const double *ray_dev, const float *filt_depth_dev, int64_t N, int P, int tail,
                                   int X, int Y, int Z, void *workspace_dev, size_t workspace_bytes, void *stream)
{
    if (!ray_dev || !filt_depth_dev || !frame_args_ok(N, P, tail) || X <= 0 || Y <= 0 || Z <= 0) return OJDF_ERR_BADARG;
    if (!workspace_dev) return OJDF_ERR_WORKSPACE;
    const long long NT = (long long)N * tail, entries = NT * 8;
    if ((long long)X * Y * Z >= 0xFFFFFFFFll || entries >= 0x7FFFFFFFll - kSegment) return OJDF_ERR_TOOLARGE;
    if (N == 0) return 0;
    Workspace ws;
    int rc;
    if ((rc = bind_workspace(ws, workspace_dev, workspace_bytes, entries)) != 0) return rc;
    Source src = {ray_dev, filt_depth_dev, nullptr, nullptr, nullptr, nullptr, NT, P, tail, 0.0f, X, Y, Z};
    return run_plan<true>(src, ws, (cudaStream_t)stream);
}

extern "C" int ojdf_integrate_apply(const float *est_dev, int64_t N, int P, int tail, float clamp_value, void *tsdf_dev,
                                    void *wvol_dev, int X, int Y, int Z, const uint8_t *pix_ids_dev,
                                    const float *pix_scores_dev, uint8_t *ids_vol_dev, void *scores_vol_dev,
                                    int do_semantics, void *workspace_dev, size_t workspace_bytes, void *stream)
{
    if (!est_dev || !frame_args_ok(N, P, tail)) return OJDF_ERR_BADARG;
    const long long NT = (long long)N * tail, entries = NT * 8;
    int rc = check_common(tsdf_dev, wvol_dev, X, Y, Z, pix_ids_dev, pix_scores_dev, ids_vol_dev, scores_vol_dev,
                          do_semantics, workspace_dev, entries);
    if (rc) return rc;
    if (N == 0) return 0;
    Workspace ws;
    if ((rc = bind_workspace(ws, workspace_dev, workspace_bytes, entries)) != 0) return rc;
    Apply a = {(__half *)tsdf_dev, (__half *)wvol_dev, ids_vol_dev, (__half *)scores_vol_dev, est_dev, pix_ids_dev,
               pix_scores_dev, (uint32_t)tail * 8u, P, 1, do_semantics, clamp_value};
    return run_apply(NT, a, ws, (cudaStream_t)stream);
}

extern "C" int ojdf_integrate(const double *ray_dev, const float *filt_depth_dev, const float *est_dev, int64_t N,
                              int P, int tail, float clamp_value, void *tsdf_dev, void *wvol_dev, int X, int Y, int Z,
                              const uint8_t *pix_ids_dev, const float *pix_scores_dev, uint8_t *ids_vol_dev,
                              void *scores_vol_dev, int do_semantics, void *workspace_dev, size_t workspace_bytes,
                              void *stream)
{
    if (!est_dev || !tsdf_dev || !wvol_dev) return OJDF_ERR_BADARG;
    int rc = ojdf_integrate_plan(ray_dev, filt_depth_dev, N, P, tail, X, Y, Z, workspace_dev, workspace_bytes, stream);
    if (rc) return rc;
    return ojdf_integrate_apply(est_dev, N, P, tail, clamp_value, tsdf_dev, wvol_dev, X, Y, Z, pix_ids_dev, pix_scores_dev,
                                ids_vol_dev, scores_vol_dev, do_semantics, workspace_dev, workspace_bytes, stream);
}

extern "C" int ojdf_integrate_updates(const float *values_dev, const int64_t *idx_dev, const double *w_dev, int64_t M1,
                                      void *tsdf_dev, void *wvol_dev, int X, int Y, int Z, const uint8_t *ids_dev,
                                      const float *scores_dev, uint8_t *ids_vol_dev, void *scores_vol_dev,
                                      int do_semantics, void *workspace_dev, size_t workspace_bytes, void *stream)
{
    if (!values_dev || !idx_dev || !w_dev || M1 < 0) return OJDF_ERR_BADARG;
    const long long entries = (long long)M1 * 8;
    int rc = check_common(tsdf_dev, wvol_dev, X, Y, Z, ids_dev, scores_dev, ids_vol_dev, scores_vol_dev, do_semantics,
                          workspace_dev, entries);
    if (rc) return rc;
    if (M1 == 0) return 0;
    Workspace ws;
    if ((rc = bind_workspace(ws, workspace_dev, workspace_bytes, entries)) != 0) return rc;
    Source src = {nullptr, nullptr, nullptr, values_dev, (const long long *)idx_dev, w_dev, M1, 0, 0, 0.0f, X, Y, Z};
    if ((rc = run_plan<false>(src, ws, (cudaStream_t)stream)) != 0) return rc;
    Apply a = {(__half *)tsdf_dev, (__half *)wvol_dev, ids_vol_dev, (__half *)scores_vol_dev, values_dev, ids_dev, scores_dev,
               8u, 0, 0, do_semantics, 0.0f};
    return run_apply(M1, a, ws, (cudaStream_t)stream);
}
