// Integrator hot path (modules/pipeline.py:137-171 + modules/integrator.py:15-126) as two
// sm_100a kernels over a caller-provided workspace -- no G^3 scratch, no sort:
//
//   scatter kernel  : one thread per (ray,sample).  Recomputes the 8 corners / weights from
//                     the extractor's per-ray record, and for every in-grid corner pushes a
//                     16-byte node {next, w, w*v} onto the touched voxel's list.  Voxels are
//                     found through an open-addressing hash keyed by the linear voxel index;
//                     the node's array position IS its entry number e = (ray*T+sample)*8+corner,
//                     so stores are 128 B per thread and fully coalesced.  The thread that
//                     first claims a voxel appends its slot to the block's touched list.
//   finalize kernel : one thread per touched voxel.  Walks the list, orders the entries by
//                     e (the reference's CPU index_add_ order), sums
//                     w and w*v sequentially in fp32, applies the running-mean update with
//                     fp16 round-to-nearest stores, resolves the semantic "last writer wins"
//                     (highest e, App. A.5) and returns the hash slot to its empty state.
//
// The result is deterministic and bit-identical to the single-threaded reference; there
// are no floating-point atomics anywhere.
#include "ojdf_internal.h"

namespace ojdf {

constexpr uint32_t kEmpty = 0xFFFFFFFFu;
constexpr int kScatterThreads = 256;
constexpr int kSegment = kScatterThreads * 8;       // worst case: every corner claims a new voxel
constexpr int kChunk = 32;                           // entries ordered per pass in finalize

struct Workspace {
    uint2 *table;        // {key, head} per slot, all-ones when idle
    uint4 *nodes;        // {next, w bits, (w*v) bits, unused} indexed by entry number
    uint32_t *list;      // per-block segments of touched slots
    uint32_t *count;     // touched slots per scatter block
    int log2_slots;
    long long max_entries;
};

static inline int log2_slots_for(long long max_entries)
{
    long long want = max_entries + max_entries / 2 + 1024;      // load factor <= 2/3 even if every entry is unique
    int l = 10;
    while ((1ll << l) < want) ++l;
    return l;
}

static size_t align256(size_t x) { return (x + 255) & ~(size_t)255; }

// Lay the workspace out for `cap` entries; returns the bytes needed.  base may be null
// (sizing only).
static size_t carve(Workspace &ws, void *base, long long cap)
{
    const int l = log2_slots_for(cap);
    const long long blocks = cap / kSegment + 1;
    const uintptr_t b = (uintptr_t)base;
    size_t off = 0;
    ws.table = (uint2 *)(b + off); off += align256(sizeof(uint2) << l);
    ws.nodes = (uint4 *)(b + off); off += align256(sizeof(uint4) * (size_t)cap);
    ws.list = (uint32_t *)(b + off); off += align256(sizeof(uint32_t) * (size_t)blocks * kSegment);
    ws.count = (uint32_t *)(b + off); off += align256(sizeof(uint32_t) * (size_t)blocks);
    ws.log2_slots = l;
    ws.max_entries = cap;
    return off;
}

static long long round_up_entries(long long e) { return (e + kSegment - 1) / kSegment * kSegment; }

// The layout is a function of the workspace SIZE only, so that calls with different frame
// sizes on one workspace all see the same (idle) hash table.
static long long capacity_for_bytes(size_t bytes)
{
    long long lo = 0, hi = (1ll << 31) / kSegment;
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

__device__ __forceinline__ uint32_t hash_slot(uint32_t key, int log2_slots)
{
    return (key * 2654435761u) >> (32 - log2_slots);
}

// Find or claim the slot of voxel `key`; returns the slot, sets is_new when this thread claimed it.
__device__ __forceinline__ uint32_t table_insert(uint2 *table, int log2_slots, uint32_t key, bool &is_new)
{
    const uint32_t mask = (1u << log2_slots) - 1u;
    uint32_t s = hash_slot(key, log2_slots);
    is_new = false;
    while (true) {
        uint32_t cur = *reinterpret_cast<volatile uint32_t *>(&table[s].x);
        if (cur == kEmpty) {
            cur = atomicCAS(&table[s].x, kEmpty, key);
            if (cur == kEmpty) { is_new = true; return s; }
        }
        if (cur == key) return s;
        s = (s + 1) & mask;
    }
}

// Push entry e onto the voxel's list and record the block's newly claimed slots.
__device__ __forceinline__ void push_entry(const Workspace &ws, uint32_t key, uint32_t e, float w, float u,
                                           uint32_t *s_count, uint32_t *seg)
{
    bool is_new;
    const uint32_t s = table_insert(ws.table, ws.log2_slots, key, is_new);
    const uint32_t prev = atomicExch(&ws.table[s].y, e);
    ws.nodes[e] = make_uint4(prev, __float_as_uint(w), __float_as_uint(u), 0u);
    if (is_new) seg[atomicAdd(s_count, 1u)] = s;
}

__global__ void __launch_bounds__(kScatterThreads)
scatter_frame_kernel(const double *__restrict__ ray, const float *__restrict__ filt, const float *__restrict__ est,
                     long long NT, int P, int T, float clampv, int X, int Y, int Z, Workspace ws)
{
    __shared__ uint32_t s_count;
    if (threadIdx.x == 0) s_count = 0;
    __syncthreads();
    uint32_t *seg = ws.list + (size_t)blockIdx.x * kSegment;
    const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (t < NT) {
        const long long n = t / T;
        const int k = (int)(t - n * T);
        if (filt[n] != 0.0f) {                                         // modules/pipeline.py:143-146
            const int i = k - P / 2;
            const double2 *rp = reinterpret_cast<const double2 *>(ray + 6 * n);
            const double2 r0 = __ldg(rp), r1 = __ldg(rp + 1), r2 = __ldg(rp + 2);
            const Axis ax = axis_setup(ray_sample(r0.x, r1.y, i));
            const Axis ay = axis_setup(ray_sample(r0.y, r2.x, i));
            const Axis az = axis_setup(ray_sample(r1.x, r2.y, i));
            float val = est[n * P + k];
            val = val < -clampv ? -clampv : (val > clampv ? clampv : val);   // torch.clamp, pipeline.py:157-159
#pragma unroll
            for (int c = 0; c < 8; ++c) {
                long long ix, iy, iz;
                if (!corner_index(ax, ay, az, c, X, Y, Z, ix, iy, iz)) continue;
                const float w = (float)corner_weight(ax, ay, az, c);        // .float(), integrator.py:45
                const float u = __fmul_rn(w, val);                          // integrator.py:55
                const uint32_t key = (uint32_t)((ix * Y + iy) * (long long)Z + iz);
                push_entry(ws, key, (uint32_t)(t * 8 + c), w, u, &s_count, seg);
            }
        }
    }
    __syncthreads();
    if (threadIdx.x == 0) ws.count[blockIdx.x] = s_count;
}

__global__ void __launch_bounds__(kScatterThreads)
scatter_updates_kernel(const float *__restrict__ values, const long long *__restrict__ idx, const double *__restrict__ wts,
                       long long M1, int X, int Y, int Z, Workspace ws)
{
    __shared__ uint32_t s_count;
    if (threadIdx.x == 0) s_count = 0;
    __syncthreads();
    uint32_t *seg = ws.list + (size_t)blockIdx.x * kSegment;
    const long long m = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (m < M1) {
        const float val = values[m];
#pragma unroll
        for (int c = 0; c < 8; ++c) {
            const long long *ip = idx + (m * 8 + c) * 3;
            const long long ix = ip[0], iy = ip[1], iz = ip[2];
            if (!(ix >= 0 && ix < X && iy >= 0 && iy < Y && iz >= 0 && iz < Z)) continue;
            const float w = (float)wts[m * 8 + c];
            const float u = __fmul_rn(w, val);
            const uint32_t key = (uint32_t)((ix * Y + iy) * (long long)Z + iz);
            push_entry(ws, key, (uint32_t)(m * 8 + c), w, u, &s_count, seg);
        }
    }
    __syncthreads();
    if (threadIdx.x == 0) ws.count[blockIdx.x] = s_count;
}

// sem_div: entries per semantic record (T*8 per pixel in the frame form, 8 per sample in the
// updates form), so entry e reads label/score number e / sem_div.
__global__ void __launch_bounds__(kScatterThreads)
finalize_kernel(Workspace ws, __half *__restrict__ tsdf, __half *__restrict__ wvol, uint8_t *__restrict__ ids_vol,
                __half *__restrict__ scores_vol, const uint8_t *__restrict__ sem_ids,
                const float *__restrict__ sem_scores, uint32_t sem_div, int do_sem)
{
    const uint32_t cnt = ws.count[blockIdx.x];
    const uint32_t *seg = ws.list + (size_t)blockIdx.x * kSegment;
    for (uint32_t j = threadIdx.x; j < cnt; j += blockDim.x) {
        const uint32_t s = seg[j];
        const uint2 slot = ws.table[s];
        ws.table[s] = make_uint2(kEmpty, kEmpty);                 // slot back to idle for the next frame
        const uint32_t key = slot.x;

        uint8_t id_old = 0;
        float sc_old = 0.0f;
        if (do_sem) { id_old = ids_vol[key]; sc_old = __half2float(scores_vol[key]); }

        float W = 0.0f, U = 0.0f;
        long long last = -1;                                       // highest entry already consumed
        uint32_t e_label = kEmpty;                                 // highest entry whose label differs from the stored one
        uint32_t be[kChunk];
        float bw[kChunk], bu[kChunk];
        bool more = true;
        while (more) {
            more = false;
            int nb = 0;
            for (uint32_t e = slot.y; e != kEmpty;) {
                const uint4 nd = ws.nodes[e];
                const uint32_t cur = e;
                e = nd.x;
                if ((long long)cur <= last) continue;
                int pos;
                if (nb < kChunk) pos = nb++;
                else { more = true; if (cur > be[kChunk - 1]) continue; pos = kChunk - 1; }
                while (pos > 0 && be[pos - 1] > cur) { be[pos] = be[pos - 1]; bw[pos] = bw[pos - 1]; bu[pos] = bu[pos - 1]; --pos; }
                be[pos] = cur; bw[pos] = __uint_as_float(nd.y); bu[pos] = __uint_as_float(nd.z);
            }
            for (int q = 0; q < nb; ++q) {                         // ascending entry order
                W = __fadd_rn(W, bw[q]);
                U = __fadd_rn(U, bu[q]);
                if (do_sem && sem_ids[be[q] / sem_div] != id_old) e_label = be[q];
            }
            last = be[nb - 1];
        }
        const float wo = __half2float(wvol[key]), vo = __half2float(tsdf[key]);
        const float wn = __fadd_rn(wo, W);
        wvol[key] = __float2half_rn(wn);                                                   // integrator.py:77-78
        tsdf[key] = __float2half_rn(__fdiv_rn(__fadd_rn(__fmul_rn(wo, vo), U), wn));      // integrator.py:82-83
        if (do_sem) {
            const float s_last = sem_scores[(uint32_t)last / sem_div];                     // highest entry wins
            scores_vol[key] = __float2half_rn(s_last > sc_old ? s_last : sc_old);          // integrator.py:112-113,124
            if (e_label != kEmpty) {
                const uint32_t r = e_label / sem_div;
                ids_vol[key] = sem_scores[r] > sc_old ? sem_ids[r] : id_old;                // integrator.py:115-116,123
            }
        }
    }
}

}  // namespace ojdf

using namespace ojdf;

extern "C" size_t ojdf_integrate_workspace_bytes(int64_t max_entries)
{
    if (max_entries <= 0) return 0;
    if (max_entries >= 0x7FFFFFFFll) return 0;
    Workspace ws;
    return carve(ws, nullptr, round_up_entries(max_entries));
}

extern "C" int ojdf_integrate_workspace_init(void *workspace_dev, size_t workspace_bytes, void *stream)
{
    if (!workspace_dev || workspace_bytes == 0) return OJDF_ERR_WORKSPACE;
    const cudaError_t e = cudaMemsetAsync(workspace_dev, 0xFF, workspace_bytes, (cudaStream_t)stream);
    return e == cudaSuccess ? 0 : (int)e;
}

static int check_common(const void *tsdf, const void *wvol, int X, int Y, int Z, const uint8_t *ids, const float *scores,
                        const uint8_t *ids_vol, const void *scores_vol, int do_sem, const void *ws, long long entries)
{
    if (!tsdf || !wvol || X <= 0 || Y <= 0 || Z <= 0) return OJDF_ERR_BADARG;
    if (do_sem && (!ids || !scores || !ids_vol || !scores_vol)) return OJDF_ERR_BADARG;
    if (!ws) return OJDF_ERR_WORKSPACE;
    if ((long long)X * Y * Z >= 0xFFFFFFFFll || entries >= 0x7FFFFFFFll) return OJDF_ERR_TOOLARGE;
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
    cudaStream_t s = (cudaStream_t)stream;
    const unsigned blocks = (unsigned)((NT + kScatterThreads - 1) / kScatterThreads);
    scatter_frame_kernel<<<blocks, kScatterThreads, 0, s>>>(ray_dev, filt_depth_dev, est_dev, NT, P, tail, clamp_value,
                                                           X, Y, Z, ws);
    finalize_kernel<<<blocks, kScatterThreads, 0, s>>>(ws, (__half *)tsdf_dev, (__half *)wvol_dev, ids_vol_dev,
                                                      (__half *)scores_vol_dev, pix_ids_dev, pix_scores_dev,
                                                      (uint32_t)tail * 8u, do_semantics);
    return launched(2);
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
    cudaStream_t s = (cudaStream_t)stream;
    const unsigned blocks = (unsigned)((M1 + kScatterThreads - 1) / kScatterThreads);
    scatter_updates_kernel<<<blocks, kScatterThreads, 0, s>>>(values_dev, (const long long *)idx_dev, w_dev, M1, X, Y, Z, ws);
    finalize_kernel<<<blocks, kScatterThreads, 0, s>>>(ws, (__half *)tsdf_dev, (__half *)wvol_dev, ids_vol_dev,
                                                      (__half *)scores_vol_dev, ids_dev, scores_dev, 8u, do_semantics);
    return launched(2);
}
