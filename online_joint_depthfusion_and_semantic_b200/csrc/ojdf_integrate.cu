// Integrator hot path (modules/pipeline.py:137-171 + modules/integrator.py:15-126) on sm_100a.
//
// The reference sums, per voxel, the contributions w and w*v of every (ray,sample,corner)
// entry that lands in it, in ascending entry order e (CPU index_add_ order),
// in fp32.  fp32 addition is not associative and the result is stored as fp16, so bit-exact
// parity needs exactly that order.  The kernels below therefore do a counting sort of the
// entries by voxel followed by a per-voxel ordering by e -- no floating-point atomics, no G^3
// scratch, no global sort:
//
//   count    : one thread per (ray,sample): recompute the 8 corners from the extractor's per-ray
//              record, find/claim the voxel's slot in an open-addressing hash (key = linear voxel
//              index), bump its entry count, remember the slot per entry (coalesced 32 B/thread);
//              the thread that claims a voxel appends the slot to its block's touched list.
//   offsets  : one thread per touched voxel: block scan of the counts + one atomic per block gives
//              every voxel a contiguous segment; voxels with more than 32 entries are queued for
//              the cooperative path.
//   place    : one thread per (ray,sample) again: write {e, w, w*v} into the voxel's segment at a
//              cursor position (arrival order).
//   finalize : short voxels (<= 32 entries): one thread orders its segment by e in registers/local
//              memory; long voxels: a warp (<= 2048 entries) or a whole block sorts the segment in
//              place with an all-ascending bitonic network (works for any length, no padding).
//              Then the fp32 sums in ascending e, the running-mean update with fp16
//              round-to-nearest stores (integrator.py:77-88), the semantic "highest entry wins"
//              update (integrator.py:90-124, App. A.5), and the slot goes back to idle.
//
// The result is deterministic and bit-identical to the single-threaded reference for every list
// length (near-camera frames put >10^4 entries into one voxel; see tests).
#include "ojdf_internal.h"

namespace ojdf {

constexpr int kThreads = 256;
constexpr int kSegment = kThreads * 8;        // entries per count/place block == max voxels a block can claim
constexpr int kShort = 32;                    // entries ordered by a single thread
constexpr int kWarpMax = 2048;                // entries ordered by one warp; above: one block
constexpr uint32_t kNoSlot = 0xFFFFFFFFu;
constexpr int kCoopBlocks = 148 * 2;          // persistent blocks draining the long-voxel queues

struct Workspace {
    uint2 *table;          // {key + 1 (0 = empty), count -> cursor}; all zero when idle
    uint32_t *eslot;       // slot of every entry (kNoSlot = out of grid / masked)
    uint32_t *list;        // touched slots, per-block segments of kSegment
    uint32_t *seg_off;     // segment offset of touched voxel (same indexing as list)
    uint32_t *seg_len;     // entry count of touched voxel
    uint4 *seg;            // {e, w bits, (w*v) bits, 0} records, grouped by voxel
    uint32_t *queue_warp;  // touched-list positions of voxels with kShort < len <= kWarpMax
    uint32_t *queue_block; // ... with len > kWarpMax
    uint32_t *count;       // touched voxels per block
    uint32_t *ctrl;        // [0] segment bump cursor, [1] warp queue length, [2] block queue length; zero when idle
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
    ws.table = (uint2 *)take(sizeof(uint2) << l);
    const size_t idle_bytes = off;                 // [0, idle_bytes) must be zero between calls
    ws.eslot = (uint32_t *)take(4 * (size_t)cap);
    ws.list = (uint32_t *)take(4 * (size_t)blocks * kSegment);
    ws.seg_off = (uint32_t *)take(4 * (size_t)blocks * kSegment);
    ws.seg_len = (uint32_t *)take(4 * (size_t)blocks * kSegment);
    ws.seg = (uint4 *)take(16 * (size_t)cap);
    ws.queue_warp = (uint32_t *)take(4 * ((size_t)cap / kShort + 1));
    ws.queue_block = (uint32_t *)take(4 * ((size_t)cap / kWarpMax + 1));
    ws.count = (uint32_t *)take(4 * (size_t)blocks);
    ws.log2_slots = l;
    ws.slots_mask = (uint32_t)((1ull << l) - 1);
    (void)idle_bytes;
    return off;
}

static size_t idle_prefix_bytes(long long cap) { return align256(256) + align256(sizeof(uint2) << log2_slots_for(cap)); }

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

// Find or claim the slot of voxel `key`; is_new is set for the one thread that claimed it.
__device__ __forceinline__ uint32_t table_insert(const Workspace &ws, uint32_t key, bool &is_new)
{
    const uint32_t k1 = key + 1u;
    uint32_t s = hash_slot(key, ws.log2_slots);
    is_new = false;
    while (true) {
        uint32_t cur = *reinterpret_cast<volatile uint32_t *>(&ws.table[s].x);
        if (cur == 0u) {
            cur = atomicCAS(&ws.table[s].x, 0u, k1);
            if (cur == 0u) { is_new = true; return s; }
        }
        if (cur == k1) return s;
        s = (s + 1) & ws.slots_mask;
    }
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
        const float v = src.est[n * src.P + k], c = src.clampv;
        val = v < -c ? -c : (v > c ? c : v);                             // torch.clamp, pipeline.py:157-159
    } else {
        sample_updates(src.idx, src.wts, t, src.X, src.Y, src.Z, s);
        val = src.values[t];
    }
    return true;
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
        uint32_t slots[8];
#pragma unroll
        for (int c = 0; c < 8; ++c) slots[c] = kNoSlot;
        if (load_sample<FRAME>(src, t, s, val)) {
#pragma unroll
            for (int c = 0; c < 8; ++c) {
                if (s.key[c] == kNoSlot) continue;
                bool is_new;
                const uint32_t sl = table_insert(ws, s.key[c], is_new);
                atomicAdd(&ws.table[sl].y, 1u);
                slots[c] = sl;
                if (is_new) seg_list[atomicAdd(&s_count, 1u)] = sl;
            }
        }
        uint4 *o = reinterpret_cast<uint4 *>(ws.eslot + t * 8);
        o[0] = make_uint4(slots[0], slots[1], slots[2], slots[3]);
        o[1] = make_uint4(slots[4], slots[5], slots[6], slots[7]);
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
    uint32_t len[kSegment / kThreads], slot[kSegment / kThreads], mine = 0;
#pragma unroll
    for (int r = 0; r < kSegment / kThreads; ++r) {                    // thread owns 8 consecutive list positions
        const uint32_t j = threadIdx.x * (kSegment / kThreads) + r;
        len[r] = 0;
        if (j < cnt) { slot[r] = ws.list[seg0 + j]; len[r] = ws.table[slot[r]].y; }
        mine += len[r];
    }
    uint32_t total;
    uint32_t pre = block_exclusive_scan(mine, total, s_warp);
    if (threadIdx.x == 0) s_base = total ? atomicAdd(&ws.ctrl[0], total) : 0u;
    __syncthreads();
    pre += s_base;
    const uint32_t lane = threadIdx.x & 31, lt_mask = (1u << lane) - 1u;
#pragma unroll
    for (int r = 0; r < kSegment / kThreads; ++r) {
        const uint32_t j = threadIdx.x * (kSegment / kThreads) + r;
        const bool valid = j < cnt;
        if (valid) {
            ws.seg_off[seg0 + j] = pre;
            ws.seg_len[seg0 + j] = len[r];
            ws.table[slot[r]].y = pre;                                 // count becomes the placement cursor
            pre += len[r];
        }
        // queue the voxels that need a cooperative sort (warp-aggregated: one atomic per warp and class)
        const bool qb = valid && len[r] > (uint32_t)kWarpMax;
        const bool qw = valid && len[r] > (uint32_t)kShort && !qb;
        const uint32_t mw = __ballot_sync(0xffffffffu, qw), mb = __ballot_sync(0xffffffffu, qb);
        if (mw) {
            const int leader = __ffs(mw) - 1;
            uint32_t base = 0;
            if ((int)lane == leader) base = atomicAdd(&ws.ctrl[1], (uint32_t)__popc(mw));
            base = __shfl_sync(0xffffffffu, base, leader);
            if (qw) ws.queue_warp[base + __popc(mw & lt_mask)] = (uint32_t)(seg0 + j);
        }
        if (mb) {
            const int leader = __ffs(mb) - 1;
            uint32_t base = 0;
            if ((int)lane == leader) base = atomicAdd(&ws.ctrl[2], (uint32_t)__popc(mb));
            base = __shfl_sync(0xffffffffu, base, leader);
            if (qb) ws.queue_block[base + __popc(mb & lt_mask)] = (uint32_t)(seg0 + j);
        }
    }
}

template <bool FRAME>
__global__ void __launch_bounds__(kThreads)
place_kernel(Source src, Workspace ws)
{
    const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    Sample s;
    float val;
    if (!load_sample<FRAME>(src, t, s, val)) return;
    const uint4 *sp = reinterpret_cast<const uint4 *>(ws.eslot + t * 8);
    const uint4 a = sp[0], b = sp[1];
    const uint32_t slots[8] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w};
#pragma unroll
    for (int c = 0; c < 8; ++c) {
        if (slots[c] == kNoSlot) continue;
        const uint32_t pos = atomicAdd(&ws.table[slots[c]].y, 1u);
        const float u = __fmul_rn(s.w[c], val);                          // integrator.py:55, separately rounded
        ws.seg[pos] = make_uint4((uint32_t)(t * 8 + c), __float_as_uint(s.w[c]), __float_as_uint(u), 0u);
    }
}

// ---------------------------------------------------------------------------------------------
struct Volumes {
    __half *tsdf; __half *wvol; uint8_t *ids; __half *scores;
    const uint8_t *sem_ids; const float *sem_scores;   // per record: entry e reads record e / sem_div
    uint32_t sem_div; int do_sem;
};

struct Accum {
    float W, U; uint32_t e_last, e_label; uint8_t id_old; float sc_old;
};

__device__ __forceinline__ void accum_begin(Accum &a, const Volumes &v, uint32_t key)
{
    a.W = 0.0f; a.U = 0.0f; a.e_last = 0; a.e_label = kNoSlot; a.id_old = 0; a.sc_old = 0.0f;
    if (v.do_sem) { a.id_old = v.ids[key]; a.sc_old = __half2float(v.scores[key]); }
}

__device__ __forceinline__ void accum_add(Accum &a, const Volumes &v, uint32_t e, float w, float u)
{
    a.W = __fadd_rn(a.W, w);                                             // index_add_, ascending e
    a.U = __fadd_rn(a.U, u);
    a.e_last = e;
    if (v.do_sem && v.sem_ids[e / v.sem_div] != a.id_old) a.e_label = e;
}

__device__ __forceinline__ void accum_store(const Accum &a, const Volumes &v, uint32_t key)
{
    const float wo = __half2float(v.wvol[key]), vo = __half2float(v.tsdf[key]);
    const float wn = __fadd_rn(wo, a.W);
    v.wvol[key] = __float2half_rn(wn);                                                        // integrator.py:77-78
    v.tsdf[key] = __float2half_rn(__fdiv_rn(__fadd_rn(__fmul_rn(wo, vo), a.U), wn));         // integrator.py:82-83 (0/0 -> NaN kept)
    if (v.do_sem) {
        const float s_last = v.sem_scores[a.e_last / v.sem_div];                              // highest entry wins
        v.scores[key] = __float2half_rn(s_last > a.sc_old ? s_last : a.sc_old);               // integrator.py:112-113,124
        if (a.e_label != kNoSlot) {
            const uint32_t r = a.e_label / v.sem_div;
            v.ids[key] = v.sem_scores[r] > a.sc_old ? v.sem_ids[r] : a.id_old;                 // integrator.py:115-116,123
        }
    }
}

// All-ascending bitonic network over seg[0..len): comparators whose upper index is >= len are
// no-ops (virtual +inf padding), so any length sorts in place.  `tid`/`nthreads`/`sync` describe the
// cooperating group (a warp or a block).
template <typename Sync>
__device__ __forceinline__ void group_sort(uint4 *seg, uint32_t len, uint32_t tid, uint32_t nthreads, Sync sync)
{
    uint32_t P = 1;
    while (P < len) P <<= 1;
    for (uint32_t k = 2; k <= P; k <<= 1) {
        for (uint32_t i = tid; i < P; i += nthreads) {                   // mirror step: i <-> i ^ (k-1)
            const uint32_t j = i ^ (k - 1);
            if (j > i && j < len) {
                const uint4 a = seg[i], b = seg[j];
                if (a.x > b.x) { seg[i] = b; seg[j] = a; }
            }
        }
        sync();
        for (uint32_t s = k >> 2; s >= 1; s >>= 1) {                     // half-cleaners: i <-> i ^ s
            for (uint32_t i = tid; i < P; i += nthreads) {
                const uint32_t j = i ^ s;
                if (j > i && j < len) {
                    const uint4 a = seg[i], b = seg[j];
                    if (a.x > b.x) { seg[i] = b; seg[j] = a; }
                }
            }
            sync();
        }
    }
}

__device__ __forceinline__ void finalize_sorted(const Workspace &ws, const Volumes &vol, uint32_t pos)
{
    const uint32_t slot = ws.list[pos], off = ws.seg_off[pos], len = ws.seg_len[pos];
    const uint32_t key = ws.table[slot].x - 1u;
    ws.table[slot] = make_uint2(0u, 0u);                                 // slot back to idle for the next frame
    Accum a;
    accum_begin(a, vol, key);
    for (uint32_t q = 0; q < len; ++q) {
        const uint4 r = ws.seg[off + q];
        accum_add(a, vol, r.x, __uint_as_float(r.y), __uint_as_float(r.z));
    }
    accum_store(a, vol, key);
}

// grid = kCoopBlocks persistent blocks that drain the two long-voxel queues (scheduled first: they
// are the long poles) + the count/place blocks (short voxels, one thread each).
__global__ void __launch_bounds__(kThreads)
finalize_kernel(Workspace ws, Volumes vol)
{
    if (blockIdx.x >= (uint32_t)kCoopBlocks) {
        const uint32_t sb = blockIdx.x - kCoopBlocks;
        const uint32_t cnt = ws.count[sb];
        const size_t seg0 = (size_t)sb * kSegment;
        for (uint32_t j = threadIdx.x; j < cnt; j += blockDim.x) {
            const uint32_t len = ws.seg_len[seg0 + j];
            if (len > (uint32_t)kShort) continue;                        // handled cooperatively below
            const uint32_t slot = ws.list[seg0 + j], off = ws.seg_off[seg0 + j];
            const uint32_t key = ws.table[slot].x - 1u;
            ws.table[slot] = make_uint2(0u, 0u);
            uint32_t be[kShort];
            float bw[kShort], bu[kShort];
            for (uint32_t q = 0; q < len; ++q) {                         // insertion sort by e while loading
                const uint4 r = ws.seg[off + q];
                int p = (int)q;
                while (p > 0 && be[p - 1] > r.x) { be[p] = be[p - 1]; bw[p] = bw[p - 1]; bu[p] = bu[p - 1]; --p; }
                be[p] = r.x; bw[p] = __uint_as_float(r.y); bu[p] = __uint_as_float(r.z);
            }
            Accum a;
            accum_begin(a, vol, key);
            for (uint32_t q = 0; q < len; ++q) accum_add(a, vol, be[q], bw[q], bu[q]);
            accum_store(a, vol, key);
        }
        return;
    }
    const uint32_t cb = blockIdx.x;
    // block queue first (rare, longest), then the warp queue
    const uint32_t nblock = ws.ctrl[2], nwarp = ws.ctrl[1];
    for (uint32_t q = cb; q < nblock; q += kCoopBlocks) {
        const uint32_t pos = ws.queue_block[q];
        group_sort(ws.seg + ws.seg_off[pos], ws.seg_len[pos], threadIdx.x, blockDim.x, [] { __syncthreads(); });
        if (threadIdx.x == 0) finalize_sorted(ws, vol, pos);
        __syncthreads();
    }
    const uint32_t lane = threadIdx.x & 31, warps_total = kCoopBlocks * (kThreads / 32);
    for (uint32_t q = cb * (kThreads / 32) + (threadIdx.x >> 5); q < nwarp; q += warps_total) {
        const uint32_t pos = ws.queue_warp[q];
        group_sort(ws.seg + ws.seg_off[pos], ws.seg_len[pos], lane, 32u, [] { __syncwarp(); });
        if (lane == 0) finalize_sorted(ws, vol, pos);
        __syncwarp();
    }
}

// Runs after finalize (stream order): control words back to idle.
__global__ void reset_ctrl_kernel(Workspace ws) { if (threadIdx.x < 4) ws.ctrl[threadIdx.x] = 0u; }

template <bool FRAME>
static int run(const Source &src, const Volumes &vol, Workspace &ws, cudaStream_t s)
{
    const unsigned blocks = (unsigned)((src.items + kThreads - 1) / kThreads);
    count_kernel<FRAME><<<blocks, kThreads, 0, s>>>(src, ws);
    offsets_kernel<<<blocks, kThreads, 0, s>>>(ws);
    place_kernel<FRAME><<<blocks, kThreads, 0, s>>>(src, ws);
    finalize_kernel<<<blocks + kCoopBlocks, kThreads, 0, s>>>(ws, vol);
    reset_ctrl_kernel<<<1, 32, 0, s>>>(ws);
    return launched(5);
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

extern "C" int ojdf_integrate(const double *ray_dev, const float *filt_depth_dev, const float *est_dev, int64_t N,
                              int P, int tail, float clamp_value, void *tsdf_dev, void *wvol_dev, int X, int Y, int Z,
                              const uint8_t *pix_ids_dev, const float *pix_scores_dev, uint8_t *ids_vol_dev,
                              void *scores_vol_dev, int do_semantics, void *workspace_dev, size_t workspace_bytes,
                              void *stream)
{
    if (!ray_dev || !filt_depth_dev || !est_dev || N < 0 || P < 1 || P > 33 || !(P & 1) || tail < 1 || tail > P)
        return OJDF_ERR_BADARG;
    const long long NT = (long long)N * tail, entries = NT * 8;
    int rc = check_common(tsdf_dev, wvol_dev, X, Y, Z, pix_ids_dev, pix_scores_dev, ids_vol_dev, scores_vol_dev,
                          do_semantics, workspace_dev, entries);
    if (rc) return rc;
    if (N == 0) return 0;
    Workspace ws;
    if ((rc = bind_workspace(ws, workspace_dev, workspace_bytes, entries)) != 0) return rc;
    Source src = {ray_dev, filt_depth_dev, est_dev, nullptr, nullptr, nullptr, NT, P, tail, clamp_value, X, Y, Z};
    Volumes vol = {(__half *)tsdf_dev, (__half *)wvol_dev, ids_vol_dev, (__half *)scores_vol_dev,
                   pix_ids_dev, pix_scores_dev, (uint32_t)tail * 8u, do_semantics};
    return run<true>(src, vol, ws, (cudaStream_t)stream);
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
    Volumes vol = {(__half *)tsdf_dev, (__half *)wvol_dev, ids_vol_dev, (__half *)scores_vol_dev,
                   ids_dev, scores_dev, 8u, do_semantics};
    return run<false>(src, vol, ws, (cudaStream_t)stream);
}
