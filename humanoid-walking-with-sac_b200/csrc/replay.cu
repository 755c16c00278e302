// HBM-resident replay: uniform ring (deque semantics) and prioritized buffer.
//
// Reference: replay_buffer.py:5-22 (ReplayBuffer) and :25-90 (PrioritizedReplayBuffer).
// Ring row = one transition [s (obs) | s2 (obs) | a (act) | r | d], fp32, padded to 16 B: a sampled row is one
// contiguous ~2.9 KB read (Humanoid), gathered by one warp with coalesced loads.
//
// PER sampling reproduces np.random.choice(len, B, p=probs) bit-for-bit downstream of the p**alpha table:
//   per_sum_*    float32 total with numpy's pairwise-summation tree (exact same association)
//   per_chunk    probs = p_alpha / total (float32), float64 chunk sums, count of "fine" elements
//   per_carry    exclusive scan of the chunk sums
//   per_search   per-sample inverse-CDF search + provable ambiguity test (see DESIGN.md "PER exactness")
//   per_exact    sequential float64 cumsum (numpy's own order) only for samples the test could not certify
//   per_finish   IS weights (N p)^-beta / max, logical indices, ring slots for the update's gather
#include <cmath>
#include <cstring>
#include <algorithm>

#include "handle.h"

namespace sacb {

// Programmatic dependent launch for the chain of small dependent kernels of one sample() call: every kernel first waits for
// its predecessor's results, then lets its successor become resident, so launch latencies overlap instead of adding up.
#define SACB_PDL_ENTER()                                                  \
    do {                                                                  \
        asm volatile("griddepcontrol.wait;" ::: "memory");                \
        asm volatile("griddepcontrol.launch_dependents;" ::: "memory");   \
    } while (0)

template <typename... KArgs, typename... Args>
static cudaError_t launch_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t stream, bool pdl, Args... args) {
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr; cfg.numAttrs = pdl ? 1 : 0;
    return cudaLaunchKernelEx(&cfg, kernel, KArgs(args)...);
}

constexpr int kChunk = 1024;          // elements per scan chunk (one warp, 32 per lane)
constexpr int kSumLeafMax = 128;      // numpy PW_BLOCKSIZE
constexpr int kSumBlockMax = 4096;    // elements handled by one CTA of per_sum_blocks
constexpr int kSumHeap = 256;         // heap slots per CTA subtree (depth <= 7)

struct PerWs {            // device workspace of one agent's PER sampler
    float *block_vals;    // heap of the top-level summation tree
    float *total;         // [1]
    double *chunk_sum;    // [n_chunks]
    double *chunk_carry;  // [n_chunks + 1] exclusive, last = cdf_last
    int *chunk_fine;      // [n_chunks]
    int *counters;        // [0] total fine, [1] flagged count, [2] exact fallbacks run
    double *cdf_exact;    // [capacity] only touched by per_exact
    double *u;            // [maxB]
    int *flagged;         // [maxB]
    int64_t *idx;         // [maxB]
    float *weights;       // [maxB]
};

// ---------------------------------------------------------------------------------------------------------------
// numpy pairwise sum tree: node (start,n) -> children (start,n2) , (start+n2, n-n2) with n2 = n/2 - (n/2)%8
// heap id 1 = root; path bits from the top give left(0)/right(1).  returns false if an ancestor was already a leaf.
// ---------------------------------------------------------------------------------------------------------------
__host__ __device__ inline bool pw_node(int64_t root_start, int64_t root_n, int limit, unsigned id, int64_t &start, int64_t &n, bool &is_leaf) {
    start = root_start; n = root_n;
#ifdef __CUDA_ARCH__
    const int depth = 31 - __clz((int)id);
#else
    const int depth = 31 - __builtin_clz(id);
#endif
    for (int b = depth - 1; b >= 0; b--) {
        if (n <= limit) return false;
        int64_t n2 = n / 2; n2 -= n2 % 8;
        if ((id >> b) & 1) { start += n2; n -= n2; } else { n = n2; }
    }
    is_leaf = n <= limit;
    return true;
}

// one numpy leaf (n <= 128) by one warp: data staged in smem, lanes 0..7 own the 8 accumulators
__device__ __forceinline__ float pw_leaf(const float *a, int n, float *sbuf, int lane) {
    for (int i = lane; i < n; i += 32) sbuf[i] = a[i];
    __syncwarp();
    float res = 0.f;
    if (n < 8) {
        if (lane == 0) for (int i = 0; i < n; i++) res = __fadd_rn(res, sbuf[i]);
    } else {
        float r = 0.f;
        const int full = n - (n % 8);
        if (lane < 8) { r = sbuf[lane]; for (int i = 8 + lane; i < full; i += 8) r = __fadd_rn(r, sbuf[i]); }
        const float r1 = __shfl_down_sync(0xffffffffu, r, 1);
        float p = __fadd_rn(r, r1);                       // lanes 0,2,4,6: r0+r1, r2+r3, r4+r5, r6+r7
        const float p2 = __shfl_down_sync(0xffffffffu, p, 2);
        float q = __fadd_rn(p, p2);                       // lanes 0,4: (r0+r1)+(r2+r3), (r4+r5)+(r6+r7)
        const float q4 = __shfl_down_sync(0xffffffffu, q, 4);
        res = __fadd_rn(q, q4);
        if (lane == 0) for (int i = full; i < n; i++) res = __fadd_rn(res, sbuf[i]);
    }
    __syncwarp();
    return res;   // valid in lane 0
}

// sum of the subtree (start,n), n <= kSumBlockMax, by one CTA (256 threads): leaves by warps, then bottom-up heap
__device__ float pw_block_sum(const float *a, int64_t start, int n, float *s_heap, unsigned char *s_state, float *s_leafbuf) {
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    // classify heap nodes: 0 = absent, 1 = leaf, 2 = internal
    for (unsigned id = 1 + tid; id < kSumHeap; id += blockDim.x) {
        int64_t s, m; bool leaf = false;
        const bool ex = pw_node(start, n, kSumLeafMax, id, s, m, leaf);
        s_state[id] = ex ? (leaf ? 1 : 2) : 0;
    }
    __syncthreads();
    for (unsigned id = 1 + warp; id < kSumHeap; id += blockDim.x / 32) {
        if (s_state[id] == 1) {
            int64_t s, m; bool leaf;
            pw_node(start, n, kSumLeafMax, id, s, m, leaf);
            const float v = pw_leaf(a + s, (int)m, s_leafbuf + warp * kSumLeafMax, lane);
            if (lane == 0) s_heap[id] = v;
        }
    }
    __syncthreads();
    for (int depth = 6; depth >= 0; depth--) {      // heap ids [2^depth, 2^(depth+1))
        for (unsigned id = (1u << depth) + tid; id < (2u << depth); id += blockDim.x)
            if (s_state[id] == 2) s_heap[id] = __fadd_rn(s_heap[2 * id], s_heap[2 * id + 1]);
        __syncthreads();
    }
    return s_heap[1];
}

// grid = top heap size; CTA id+1 = heap id of the top tree (leaf threshold kSumBlockMax)
__global__ void __launch_bounds__(256) per_sum_blocks(const float *p_alpha, int64_t n, float *top_vals) {
    SACB_PDL_ENTER();
    __shared__ float s_heap[kSumHeap];
    __shared__ unsigned char s_state[kSumHeap];
    __shared__ float s_leafbuf[8 * kSumLeafMax];
    const unsigned id = blockIdx.x + 1;
    int64_t s, m; bool leaf = false;
    if (!pw_node(0, n, kSumBlockMax, id, s, m, leaf) || !leaf) return;
    const float v = pw_block_sum(p_alpha, s, (int)m, s_heap, s_state, s_leafbuf);
    if (threadIdx.x == 0) top_vals[id] = v;
}

__global__ void __launch_bounds__(1024) per_sum_top(int64_t n, int top_depth, float *top_vals, float *total) {
    SACB_PDL_ENTER();
    for (int depth = top_depth; depth >= 0; depth--) {
        for (unsigned id = (1u << depth) + threadIdx.x; id < (2u << depth); id += blockDim.x) {
            int64_t s, m; bool leaf = false;
            if (pw_node(0, n, kSumBlockMax, id, s, m, leaf) && !leaf) top_vals[id] = __fadd_rn(top_vals[2 * id], top_vals[2 * id + 1]);
        }
        __syncthreads();
    }
    if (threadIdx.x == 0) *total = top_vals[1];
}

// ---------------------------------------------------------------------------------------------------------------
// chunked float64 scan
// ---------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ bool is_fine(double p) {      // not an integer multiple of 2^-52
    const double x = p * 4503599627370496.0;
    return x != floor(x);
}

// lane-local pass over 32 consecutive probs of a chunk; returns the lane sum (sequential order), counts fine elements
__device__ __forceinline__ double lane_pass(const float *p_alpha, int64_t base, int64_t n, float total, int &fine) {
    double acc = 0.0;
    fine = 0;
#pragma unroll 8
    for (int j = 0; j < 32; j++) {
        const int64_t i = base + j;
        if (i < n) {
            const double p = (double)__fdiv_rn(p_alpha[i], total);     // probs /= probs.sum()  (float32 divide)
            fine += is_fine(p) ? 1 : 0;
            acc = __dadd_rn(acc, p);
        }
    }
    return acc;
}

__device__ __forceinline__ double warp_excl_scan(double v, int lane, double &warp_total) {
    double incl = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const double t = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) incl = __dadd_rn(t, incl);
    }
    warp_total = __shfl_sync(0xffffffffu, incl, 31);
    const double ex = __shfl_up_sync(0xffffffffu, incl, 1);
    return lane == 0 ? 0.0 : ex;
}

__global__ void __launch_bounds__(256) per_chunk(const float *p_alpha, int64_t n, const float *total, double *chunk_sum, int *chunk_fine) {
    SACB_PDL_ENTER();
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int64_t chunk = (int64_t)blockIdx.x * 8 + warp;
    if (chunk * kChunk >= n) return;
    int fine;
    const double v = lane_pass(p_alpha, chunk * kChunk + lane * 32, n, *total, fine);
    double wt;
    warp_excl_scan(v, lane, wt);
    for (int o = 16; o > 0; o >>= 1) fine += __shfl_xor_sync(0xffffffffu, fine, o);
    if (lane == 0) { chunk_sum[chunk] = wt; chunk_fine[chunk] = fine; }
}

// single CTA: exclusive Kogge-Stone scan of the chunk sums (n_chunks <= 1024 per pass, looped with a running carry)
__global__ void __launch_bounds__(1024) per_carry(const double *chunk_sum, const int *chunk_fine, int n_chunks, double *carry, int *counters) {
    SACB_PDL_ENTER();
    __shared__ double s[1024];
    __shared__ int s_f[1024];
    __shared__ double s_run;
    __shared__ int s_frun;
    if (threadIdx.x == 0) { s_run = 0.0; s_frun = 0; }
    __syncthreads();
    for (int base = 0; base < n_chunks; base += 1024) {
        const int i = base + threadIdx.x;
        double v = i < n_chunks ? chunk_sum[i] : 0.0;
        int f = i < n_chunks ? chunk_fine[i] : 0;
        s[threadIdx.x] = v; s_f[threadIdx.x] = f;
        __syncthreads();
        for (int o = 1; o < 1024; o <<= 1) {
            double t = 0.0; int tf = 0;
            if (threadIdx.x >= o) { t = s[threadIdx.x - o]; tf = s_f[threadIdx.x - o]; }
            __syncthreads();
            if (threadIdx.x >= o) { s[threadIdx.x] = __dadd_rn(t, s[threadIdx.x]); s_f[threadIdx.x] += tf; }
            __syncthreads();
        }
        const double run = s_run; const int frun = s_frun;
        if (i < n_chunks) carry[i] = __dadd_rn(run, threadIdx.x ? s[threadIdx.x - 1] : 0.0);   // exclusive prefix
        __syncthreads();
        if (threadIdx.x == 1023) { s_run = __dadd_rn(run, s[1023]); s_frun = frun + s_f[1023]; }
        __syncthreads();
    }
    if (threadIdx.x == 0) { carry[n_chunks] = s_run; counters[0] = s_frun; counters[1] = 0; }
}

// one warp per sample
__global__ void __launch_bounds__(256) per_search(const float *p_alpha, int64_t n, const float *total, const double *carry, int n_chunks,
                                                  const double *u, int B, const int *counters, int64_t *idx_out, int *flagged, int *flag_count) {
    SACB_PDL_ENTER();
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int j = blockIdx.x * 8 + warp;
    if (j >= B) return;
    const double last = carry[n_chunks];
    const double uu = u[j];
    // provable bound on |cdf_parallel - cdf_sequential| / last  (DESIGN.md): zero when no element has bits below 2^-52
    const int F = counters[0];
    const double window = F == 0 ? 0.0 : ((128.0 * (double)F + 64.0) * 2.220446049250313e-16) / last * 4.0 + 4.440892098500626e-16;
    // chunk: last c with carry[c]/last <= u
    int lo = 0, hi = n_chunks - 1;
    while (lo < hi) {
        const int mid = (lo + hi + 1) >> 1;
        if (__ddiv_rn(carry[mid], last) <= uu) lo = mid; else hi = mid - 1;
    }
    const int c = lo;
    int fine;
    const int64_t base = (int64_t)c * kChunk + lane * 32;
    const double v = lane_pass(p_alpha, base, n, *total, fine);
    double wt;
    const double ex = warp_excl_scan(v, lane, wt);
    // second pass: running cdf inside the lane, count elements with cdf <= u, track the distance to the nearest boundary
    double acc = __dadd_rn(carry[c], ex);
    double mind = fabs(__ddiv_rn(carry[c], last) - uu);   // boundary with the previous chunk
    if (c == 0) mind = 1.0;
    int cnt = 0;
    const float tot = *total;
    for (int k = 0; k < 32; k++) {
        const int64_t i = base + k;
        if (i < n) {
            acc = __dadd_rn(acc, (double)__fdiv_rn(p_alpha[i], tot));
            const double cn = __ddiv_rn(acc, last);
            cnt += (cn <= uu) ? 1 : 0;
            mind = fmin(mind, fabs(cn - uu));
        }
    }
    for (int o = 16; o > 0; o >>= 1) {
        cnt += __shfl_xor_sync(0xffffffffu, cnt, o);
        mind = fmin(mind, __shfl_xor_sync(0xffffffffu, mind, o));
    }
    if (lane == 0) {
        int64_t idx = (int64_t)c * kChunk + cnt;
        const bool flag = (F != 0 && mind <= window) || idx >= n;
        if (idx >= n) idx = n - 1;
        idx_out[j] = idx;
        flagged[j] = flag ? 1 : 0;
        if (flag) atomicAdd(flag_count, 1);
    }
}

// numpy's own algorithm, run only when a sample could not be certified: sequential float64 cumsum, /= last, searchsorted right
__global__ void __launch_bounds__(32) per_exact(const float *p_alpha, int64_t n, const float *total, double *cdf, const double *u, int B,
                                                 const int *flagged, int64_t *idx_out, int *counters) {
    SACB_PDL_ENTER();
    if (counters[1] == 0) return;
    const int lane = threadIdx.x;
    const float tot = *total;
    if (lane == 0) {
        double acc = 0.0;
        for (int64_t i = 0; i < n; i++) { acc = __dadd_rn(acc, (double)__fdiv_rn(p_alpha[i], tot)); cdf[i] = acc; }
        counters[2] += 1;
    }
    __syncwarp();
    __threadfence_block();
    const double last = cdf[n - 1];
    for (int j = lane; j < B; j += 32) {
        if (!flagged[j]) continue;
        const double uu = u[j];
        int64_t lo = 0, hi = n;
        while (lo < hi) {
            const int64_t mid = lo + ((hi - lo) >> 1);
            if (uu < __ddiv_rn(cdf[mid], last)) hi = mid; else lo = mid + 1;
        }
        idx_out[j] = lo < n ? lo : n - 1;
    }
}

// single CTA: IS weights (replay_buffer.py:67-68), ring slots for the gather
__global__ void __launch_bounds__(1024) per_finish(const float *p_alpha, int64_t n, const float *total, const int64_t *idx, int B, float neg_beta,
                                                    float *weights, int32_t *slots, float *isw_ws, int64_t *idx_copy) {
    SACB_PDL_ENTER();
    __shared__ float s_max[32];
    float w = -INFINITY;
    const int j = threadIdx.x;
    if (j < B) {
        const float prob = __fdiv_rn(p_alpha[idx[j]], *total);
        w = powf((float)n * prob, neg_beta);
        slots[j] = (int32_t)idx[j];
        idx_copy[j] = idx[j];      // indices of the last sample (update_priorities without an index argument, TD write-back)
    }
    float m = w;
    for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
    if ((threadIdx.x & 31) == 0) s_max[threadIdx.x >> 5] = m;
    __syncthreads();
    if (threadIdx.x < 32) {
        m = s_max[threadIdx.x];
        for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
        if (threadIdx.x == 0) s_max[0] = m;
    }
    __syncthreads();
    if (j < B) {
        const float wn = __fdiv_rn(w, s_max[0]);
        weights[j] = wn;
        if (isw_ws) isw_ws[j] = wn;
    }
}

__global__ void per_uniform_draw(double *u, int B, uint64_t seed, uint64_t counter) {
    SACB_PDL_ENTER();
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= B) return;
    uint64_t x = seed ^ (counter * 0x9E3779B97F4A7C15ull + (uint64_t)j * 0xBF58476D1CE4E5B9ull);   // splitmix64
    x ^= x >> 30; x *= 0xBF58476D1CE4E5B9ull; x ^= x >> 27; x *= 0x94D049BB133111EBull; x ^= x >> 31;
    u[j] = (double)(x >> 11) * (1.0 / 9007199254740992.0);
}

// ---------------------------------------------------------------------------------------------------------------
// priorities: update (last duplicate wins), push (max over the whole capacity array)
// ---------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(1024) per_update_kernel(float *prio, float *p_alpha, const int64_t *idx, const float *td, int B, float alpha, int is_final) {
    SACB_PDL_ENTER();
    extern __shared__ int64_t s_idx[];
    for (int j = threadIdx.x; j < B; j += blockDim.x) s_idx[j] = idx[j];
    __syncthreads();
    for (int j = threadIdx.x; j < B; j += blockDim.x) {
        bool last = true;
        for (int k = j + 1; k < B; k++) if (s_idx[k] == s_idx[j]) { last = false; break; }
        if (last) {
            const float p = is_final ? td[j] : (float)((double)td[j] + 1e-6);      // priority.item() + 1e-6, stored as float32
            prio[s_idx[j]] = p;
            p_alpha[s_idx[j]] = powf(p, alpha);
        }
    }
}

__global__ void __launch_bounds__(1024) max_reduce_kernel(const float *x, int64_t n, float *block_out) {
    __shared__ float s[32];
    float m = -INFINITY;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) m = fmaxf(m, x[i]);
    for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
    if ((threadIdx.x & 31) == 0) s[threadIdx.x >> 5] = m;
    __syncthreads();
    if (threadIdx.x < 32) {
        m = s[threadIdx.x];
        for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
        if (threadIdx.x == 0) block_out[blockIdx.x] = m;
    }
}

__global__ void per_push_kernel(float *prio, float *p_alpha, const float *block_max, int n_blocks, int empty, int64_t pos, int64_t count,
                                int64_t capacity, float alpha) {
    float m = 1.0f;                                            // `if self.buffer else 1.0`
    if (!empty) { m = block_max[0]; for (int i = 1; i < n_blocks; i++) m = fmaxf(m, block_max[i]); }
    const float pa = powf(m, alpha);
    for (int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; k < count; k += (int64_t)gridDim.x * blockDim.x) {
        const int64_t slot = (pos + k) % capacity;
        prio[slot] = m; p_alpha[slot] = pa;
    }
}

__global__ void pow_alpha_kernel(const float *prio, float *p_alpha, int64_t n, float alpha) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) p_alpha[i] = powf(prio[i], alpha);
}

// ---------------------------------------------------------------------------------------------------------------
// ring gather to a dense staging buffer (host-facing sample / read-back)
// ---------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) gather_rows_kernel(const float *ring, int64_t ring_row, const int32_t *slots, int n, float *out) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int j = blockIdx.x * 8 + warp;
    if (j >= n) return;
    const float4 *src = reinterpret_cast<const float4 *>(ring + (int64_t)slots[j] * ring_row);
    float4 *dst = reinterpret_cast<float4 *>(out + (int64_t)j * ring_row);
    for (int k = lane; k < ring_row / 4; k += 32) dst[k] = __ldcs(src + k);
}

static PerWs per_ws_of(sacb_handle h, int agent) {
    PerWs w;
    char *base = (char *)h->per_ws;
    const int64_t cap = h->cfg.capacity, nch = cap / kChunk + 2, B = h->cfg.max_batch;
    auto take = [&](size_t bytes) { char *p = base; base += align_up((int64_t)bytes, 256); return p; };
    // same carve-up for every agent; agent stride computed by per_ws_bytes
    (void)agent;
    w.block_vals = (float *)take(sizeof(float) * 4096);
    w.total = (float *)take(256);
    w.chunk_sum = (double *)take(sizeof(double) * nch);
    w.chunk_carry = (double *)take(sizeof(double) * (nch + 1));
    w.chunk_fine = (int *)take(sizeof(int) * nch);
    w.counters = (int *)take(256);
    w.cdf_exact = (double *)take(sizeof(double) * cap);
    w.u = (double *)take(sizeof(double) * B);
    w.flagged = (int *)take(sizeof(int) * B);
    w.idx = (int64_t *)take(sizeof(int64_t) * B);
    w.weights = (float *)take(sizeof(float) * B);
    return w;
}
static int64_t per_ws_bytes(sacb_handle h) {
    const int64_t cap = h->cfg.capacity, nch = cap / kChunk + 2, B = h->cfg.max_batch;
    return 4096 * 4 + 256 + 8 * nch + 8 * (nch + 1) + 4 * nch + 256 + 8 * cap + 8 * B + 4 * B + 8 * B + 4 * B + 16 * 256;
}

int replay_create(sacb_handle h) {
    const sacb_config &c = h->cfg;
    h->ring_row = align_up(2 * c.obs_dim + c.act_dim + 2, 4);
    const int n = c.n_agents;
    h->r_len.assign(n, 0); h->r_pos.assign(n, 0); h->r_head.assign(n, 0); h->per_frame.assign(n, 1);   // frame starts at 1 (replay_buffer.py:31)
    const size_t ring_bytes = sizeof(float) * (size_t)h->ring_row * c.capacity * n;
    if (cudaMalloc(&h->ring, ring_bytes) != cudaSuccess) { cudaGetLastError(); return fail(SACB_ERR_NOMEM, "replay ring does not fit in device memory"); }
    h->stage_rows_cap = std::max<int64_t>(c.max_batch, 4096);
    if (cudaMalloc(&h->stage_rows, sizeof(float) * h->ring_row * h->stage_rows_cap) != cudaSuccess) return fail(SACB_ERR_NOMEM, "staging alloc failed");
    if (c.replay_kind == SACB_REPLAY_PER) {
        if (n != 1) return fail(SACB_ERR_ARG, "prioritized replay is per-agent: use n_agents = 1 per handle");
        if (c.max_batch > 1024) return fail(SACB_ERR_ARG, "prioritized replay supports max_batch <= 1024");
        if (cudaMalloc(&h->prio, sizeof(float) * c.capacity) != cudaSuccess || cudaMalloc(&h->p_alpha, sizeof(float) * c.capacity) != cudaSuccess ||
            cudaMalloc(&h->per_ws, per_ws_bytes(h)) != cudaSuccess || cudaMalloc(&h->last_idx_dev, sizeof(int64_t) * c.max_batch) != cudaSuccess)
            return fail(SACB_ERR_NOMEM, "PER tables do not fit in device memory");
        cudaMemsetAsync(h->prio, 0, sizeof(float) * c.capacity, h->stream);
        cudaMemsetAsync(h->p_alpha, 0, sizeof(float) * c.capacity, h->stream);
        cudaMemsetAsync(h->per_ws, 0, per_ws_bytes(h), h->stream);
    }
    return SACB_OK;
}

void replay_destroy(sacb_handle h) {
    cudaFree(h->ring); cudaFree(h->stage_rows); cudaFree(h->prio); cudaFree(h->p_alpha); cudaFree(h->per_ws); cudaFree(h->last_idx_dev);
}

// logical index -> physical ring slot.  uniform: deque order (j-th oldest); PER: list position
static inline int64_t physical_slot(sacb_handle h, int agent, int64_t j) {
    if (h->cfg.replay_kind == SACB_REPLAY_PER) return j;
    return (h->r_head[agent] + j) % h->cfg.capacity;
}

int replay_stage_slots(sacb_handle h, const int64_t *idx, int64_t B) {
    const int n = h->cfg.n_agents;
    int32_t *pin = reinterpret_cast<int32_t *>(h->pin);
    if ((int64_t)n * B * (int64_t)sizeof(int32_t) > h->pin_floats * (int64_t)sizeof(float)) return fail(SACB_ERR_ARG, "index set too large");
    for (int a = 0; a < n; a++)
        for (int64_t j = 0; j < B; j++) {
            const int64_t lj = idx[a * B + j];
            if (lj < 0 || lj >= h->r_len[a]) return fail(SACB_ERR_STATE, "replay index out of range");
            pin[a * B + j] = (int32_t)physical_slot(h, a, lj);
        }
    SACB_CUDA(cudaStreamSynchronize(h->stream));   // the pinned buffer is reused by every call
    for (int a = 0; a < n; a++)
        SACB_CUDA(cudaMemcpyAsync(h->slots + (int64_t)a * h->cfg.max_batch, pin + a * B, sizeof(int32_t) * B, cudaMemcpyHostToDevice, h->stream));
    SACB_CUDA(cudaStreamSynchronize(h->stream));
    return SACB_OK;
}

}  // namespace sacb
using namespace sacb;

extern "C" int sacb_stage_indices(sacb_handle h, const int64_t *idx, int64_t B, int64_t n_steps) {
    if (!h || !idx || B < 1 || B > h->cfg.max_batch || n_steps < 1) return fail(SACB_ERR_ARG, "bad argument");
    const int n = h->cfg.n_agents;
    std::vector<int32_t> slots((size_t)n_steps * n * B);
    for (int64_t s = 0; s < n_steps; s++)
        for (int a = 0; a < n; a++)
            for (int64_t j = 0; j < B; j++) {
                const int64_t lj = idx[(s * n + a) * B + j];
                if (lj < 0 || lj >= h->r_len[a]) return fail(SACB_ERR_STATE, "replay index out of range");
                slots[(s * n + a) * B + j] = (int32_t)physical_slot(h, a, lj);
            }
    if (B != h->cfg.max_batch && n > 1) return fail(SACB_ERR_ARG, "population staging needs B == max_batch");
    cudaFree(h->slots_staged);
    SACB_CUDA(cudaMalloc(&h->slots_staged, slots.size() * sizeof(int32_t)));
    SACB_CUDA(cudaMemcpy(h->slots_staged, slots.data(), slots.size() * sizeof(int32_t), cudaMemcpyHostToDevice));
    h->staged_steps = n_steps; h->staged_next = 0; h->staged_B = B;
    return SACB_OK;
}

extern "C" int64_t sacb_len(sacb_handle h, int agent) {
    if (!h || agent < 0 || agent >= h->cfg.n_agents) return -1;
    return h->r_len[agent];
}

extern "C" int sacb_clear_replay(sacb_handle h, int agent) {
    if (!h || agent < 0 || agent >= h->cfg.n_agents) return fail(SACB_ERR_ARG, "bad argument");
    h->r_len[agent] = h->r_pos[agent] = h->r_head[agent] = 0;
    if (h->prio) { cudaMemsetAsync(h->prio, 0, sizeof(float) * h->cfg.capacity, h->stream); cudaMemsetAsync(h->p_alpha, 0, sizeof(float) * h->cfg.capacity, h->stream); }
    return SACB_OK;
}

static int per_push_priorities(sacb_handle h, int64_t pos, int64_t count, bool empty) {
    PerWs w = per_ws_of(h, 0);
    const int nb = 128;
    if (!empty) max_reduce_kernel<<<nb, 1024, 0, h->stream>>>(h->prio, h->cfg.capacity, w.block_vals);
    per_push_kernel<<<(int)std::min<int64_t>(64, (count + 255) / 256), 256, 0, h->stream>>>(h->prio, h->p_alpha, w.block_vals, nb, empty ? 1 : 0, pos, count,
                                                                                          h->cfg.capacity, h->cfg.per_alpha);
    h->kernel_launches += empty ? 1 : 2;
    SACB_CUDA(cudaGetLastError());
    return SACB_OK;
}

extern "C" int sacb_push(sacb_handle h, int agent, const float *s, const float *a, const float *r, const float *s2, const float *done, int64_t n) {
    if (!h || !s || !a || !r || !s2 || !done || agent < 0 || agent >= h->cfg.n_agents || n < 0) return fail(SACB_ERR_ARG, "bad argument");
    const sacb_config &c = h->cfg;
    const int64_t cap = c.capacity, row = h->ring_row;
    float *ring = h->ring + (int64_t)agent * cap * row;
    const bool per = c.replay_kind == SACB_REPLAY_PER;
    int64_t done_n = 0;
    while (done_n < n) {
        // next write slot; a contiguous run never crosses the end of the ring
        int64_t slot;
        if (per) slot = h->r_pos[agent];
        else slot = h->r_len[agent] < cap ? (h->r_head[agent] + h->r_len[agent]) % cap : h->r_head[agent];
        const int64_t run = std::min(n - done_n, cap - slot);
        const size_t pitch = sizeof(float) * row, wo = sizeof(float) * c.obs_dim, wa = sizeof(float) * c.act_dim;
        float *dst = ring + slot * row;
        SACB_CUDA(cudaMemcpy2DAsync(dst, pitch, s + done_n * c.obs_dim, wo, wo, run, cudaMemcpyHostToDevice, h->stream));
        SACB_CUDA(cudaMemcpy2DAsync(dst + c.obs_dim, pitch, s2 + done_n * c.obs_dim, wo, wo, run, cudaMemcpyHostToDevice, h->stream));
        SACB_CUDA(cudaMemcpy2DAsync(dst + 2 * c.obs_dim, pitch, a + done_n * c.act_dim, wa, wa, run, cudaMemcpyHostToDevice, h->stream));
        SACB_CUDA(cudaMemcpy2DAsync(dst + 2 * c.obs_dim + c.act_dim, pitch, r + done_n, sizeof(float), sizeof(float), run, cudaMemcpyHostToDevice, h->stream));
        SACB_CUDA(cudaMemcpy2DAsync(dst + 2 * c.obs_dim + c.act_dim + 1, pitch, done + done_n, sizeof(float), sizeof(float), run, cudaMemcpyHostToDevice, h->stream));
        if (per) {
            int rc = per_push_priorities(h, slot, run, h->r_len[agent] == 0);
            if (rc) return rc;
            h->r_pos[agent] = (slot + run) % cap;
            h->r_len[agent] = std::min(cap, h->r_len[agent] + run);
        } else {
            const int64_t grow = std::min(run, cap - h->r_len[agent]);
            h->r_len[agent] += grow;
            h->r_head[agent] = (h->r_head[agent] + (run - grow)) % cap;      // overwritten oldest entries (deque maxlen eviction)
        }
        done_n += run;
    }
    SACB_CUDA(cudaStreamSynchronize(h->stream));     // caller may reuse its buffers
    return SACB_OK;
}

extern "C" int64_t sacb_row_floats(sacb_handle h) { return h ? h->ring_row : -1; }

extern "C" int sacb_push_rows(sacb_handle h, int agent, const float *rows, int64_t n) {
    if (!h || !rows || agent < 0 || agent >= h->cfg.n_agents || n < 0) return fail(SACB_ERR_ARG, "bad argument");
    const sacb_config &c = h->cfg;
    const int64_t cap = c.capacity, row = h->ring_row;
    float *ring = h->ring + (int64_t)agent * cap * row;
    const bool per = c.replay_kind == SACB_REPLAY_PER;
    int64_t done_n = 0;
    while (done_n < n) {
        int64_t slot;
        if (per) slot = h->r_pos[agent];
        else slot = h->r_len[agent] < cap ? (h->r_head[agent] + h->r_len[agent]) % cap : h->r_head[agent];
        const int64_t run = std::min(n - done_n, cap - slot);
        SACB_CUDA(cudaMemcpyAsync(ring + slot * row, rows + done_n * row, sizeof(float) * row * run, cudaMemcpyHostToDevice, h->stream));
        if (per) {
            int rc = per_push_priorities(h, slot, run, h->r_len[agent] == 0);
            if (rc) return rc;
            h->r_pos[agent] = (slot + run) % cap;
            h->r_len[agent] = std::min(cap, h->r_len[agent] + run);
        } else {
            const int64_t grow = std::min(run, cap - h->r_len[agent]);
            h->r_len[agent] += grow;
            h->r_head[agent] = (h->r_head[agent] + (run - grow)) % cap;
        }
        done_n += run;
    }
    SACB_CUDA(cudaStreamSynchronize(h->stream));
    return SACB_OK;
}

static int gather_to_host(sacb_handle h, int agent, const int32_t *slots_host, int64_t n, float *s, float *a, float *r, float *s2, float *done) {
    const sacb_config &c = h->cfg;
    const int64_t row = h->ring_row;
    const float *ring = h->ring + (int64_t)agent * c.capacity * row;
    int32_t *slots_dev = h->slots + (int64_t)agent * c.max_batch;
    for (int64_t off = 0; off < n; off += c.max_batch) {
        const int64_t m = std::min<int64_t>(c.max_batch, n - off);
        if (slots_host) SACB_CUDA(cudaMemcpyAsync(slots_dev, slots_host + off, sizeof(int32_t) * m, cudaMemcpyHostToDevice, h->stream));
        gather_rows_kernel<<<(int)((m + 7) / 8), 256, 0, h->stream>>>(ring, row, slots_dev, (int)m, h->stage_rows);
        h->kernel_launches++;
        const size_t pitch = sizeof(float) * row, wo = sizeof(float) * c.obs_dim, wa = sizeof(float) * c.act_dim;
        const float *src = h->stage_rows;
        if (s) SACB_CUDA(cudaMemcpy2DAsync(s + off * c.obs_dim, wo, src, pitch, wo, m, cudaMemcpyDeviceToHost, h->stream));
        if (s2) SACB_CUDA(cudaMemcpy2DAsync(s2 + off * c.obs_dim, wo, src + c.obs_dim, pitch, wo, m, cudaMemcpyDeviceToHost, h->stream));
        if (a) SACB_CUDA(cudaMemcpy2DAsync(a + off * c.act_dim, wa, src + 2 * c.obs_dim, pitch, wa, m, cudaMemcpyDeviceToHost, h->stream));
        if (r) SACB_CUDA(cudaMemcpy2DAsync(r + off, sizeof(float), src + 2 * c.obs_dim + c.act_dim, pitch, sizeof(float), m, cudaMemcpyDeviceToHost, h->stream));
        if (done) SACB_CUDA(cudaMemcpy2DAsync(done + off, sizeof(float), src + 2 * c.obs_dim + c.act_dim + 1, pitch, sizeof(float), m, cudaMemcpyDeviceToHost, h->stream));
        SACB_CUDA(cudaStreamSynchronize(h->stream));
    }
    return SACB_OK;
}

extern "C" int sacb_read_transitions(sacb_handle h, int agent, const int64_t *idx, int64_t n, float *s, float *a, float *r, float *s2, float *done) {
    if (!h || !idx || agent < 0 || agent >= h->cfg.n_agents) return fail(SACB_ERR_ARG, "bad argument");
    std::vector<int32_t> slots(n);
    for (int64_t j = 0; j < n; j++) {
        if (idx[j] < 0 || idx[j] >= h->r_len[agent]) return fail(SACB_ERR_STATE, "replay index out of range");
        slots[j] = (int32_t)physical_slot(h, agent, idx[j]);
    }
    return gather_to_host(h, agent, slots.data(), n, s, a, r, s2, done);
}

extern "C" int sacb_sample_uniform(sacb_handle h, int agent, const int64_t *idx, int64_t B, float *s, float *a, float *r, float *s2, float *done) {
    if (!h || agent < 0 || agent >= h->cfg.n_agents) return fail(SACB_ERR_ARG, "bad argument");
    if (B > h->r_len[agent]) return fail(SACB_ERR_STATE, "Sample larger than population or is negative");   // random.sample's ValueError
    if (!idx) return fail(SACB_ERR_ARG, "indices required (draw them with random.sample(range(len), B))");
    return sacb_read_transitions(h, agent, idx, B, s, a, r, s2, done);
}

// ---- PER --------------------------------------------------------------------------------------------------------
static int top_depth_of(int64_t n) {
    int d = 0;
    while (n > kSumBlockMax) { int64_t n2 = n / 2; n2 -= n2 % 8; n = n - n2; d++; }   // the right child is the larger one
    return d;
}

extern "C" int sacb_per_sample(sacb_handle h, int agent, const double *u, int64_t B, int64_t *idx_out, float *weights_out,
                               float *s, float *a, float *r, float *s2, float *done) {
    if (!h || agent != 0 || h->cfg.replay_kind != SACB_REPLAY_PER) return fail(SACB_ERR_ARG, "handle has no prioritized buffer");
    const int64_t n = h->r_len[0];
    if (n < 1) return fail(SACB_ERR_STATE, "cannot sample from an empty buffer");
    const int64_t k = std::min<int64_t>(B, n);                       // n_samples = min(batch_size, len)  (replay_buffer.py:50)
    if (k > h->cfg.max_batch) return fail(SACB_ERR_ARG, "batch size exceeds max_batch");
    PerWs w = per_ws_of(h, 0);
    cudaStream_t st = h->stream;
    const bool pdl = h->use_pdl != 0;
    const int64_t frame = h->per_frame[0];
    const double beta = std::min(1.0, (double)h->cfg.per_beta_start + (double)frame * (1.0 - (double)h->cfg.per_beta_start) / (double)h->cfg.per_beta_frames);
    h->per_frame[0] = frame + 1;
    if (u) SACB_CUDA(cudaMemcpyAsync(w.u, u, sizeof(double) * k, cudaMemcpyHostToDevice, h->stream));
    else SACB_CUDA(launch_pdl(per_uniform_draw, dim3((int)((k + 255) / 256)), dim3(256), 0, st, pdl, w.u, (int)k, h->cfg.seed, (uint64_t)frame));
    const int depth = top_depth_of(n);
    if (depth > 11) return fail(SACB_ERR_ARG, "capacity too large for the summation heap");
    const float *pa = h->p_alpha;
    SACB_CUDA(launch_pdl(per_sum_blocks, dim3((2 << depth) - 1), dim3(256), 0, st, pdl, pa, n, w.block_vals));
    SACB_CUDA(launch_pdl(per_sum_top, dim3(1), dim3(1024), 0, st, pdl, n, depth, w.block_vals, w.total));
    const int n_chunks = (int)((n + kChunk - 1) / kChunk);
    SACB_CUDA(launch_pdl(per_chunk, dim3((n_chunks + 7) / 8), dim3(256), 0, st, pdl, pa, n, (const float *)w.total, w.chunk_sum, w.chunk_fine));
    SACB_CUDA(launch_pdl(per_carry, dim3(1), dim3(1024), 0, st, pdl, (const double *)w.chunk_sum, (const int *)w.chunk_fine, n_chunks, w.chunk_carry, w.counters));
    SACB_CUDA(launch_pdl(per_search, dim3((int)((k + 7) / 8)), dim3(256), 0, st, pdl, pa, n, (const float *)w.total, (const double *)w.chunk_carry, n_chunks,
                         (const double *)w.u, (int)k, (const int *)w.counters, w.idx, w.flagged, w.counters + 1));
    SACB_CUDA(launch_pdl(per_exact, dim3(1), dim3(32), 0, st, pdl, pa, n, (const float *)w.total, w.cdf_exact, (const double *)w.u, (int)k, (const int *)w.flagged, w.idx, w.counters));
    SACB_CUDA(launch_pdl(per_finish, dim3(1), dim3(1024), 0, st, pdl, pa, n, (const float *)w.total, (const int64_t *)w.idx, (int)k, -(float)beta, w.weights, h->slots,
                         h->ws + h->L.isw, h->last_idx_dev));
    h->kernel_launches += u ? 7 : 8;
    if (idx_out) SACB_CUDA(cudaMemcpyAsync(idx_out, w.idx, sizeof(int64_t) * k, cudaMemcpyDeviceToHost, h->stream));
    if (weights_out) SACB_CUDA(cudaMemcpyAsync(weights_out, w.weights, sizeof(float) * k, cudaMemcpyDeviceToHost, h->stream));
    if (s || a || r || s2 || done) return gather_to_host(h, 0, nullptr, k, s, a, r, s2, done);
    if (idx_out || weights_out) SACB_CUDA(cudaStreamSynchronize(h->stream));
    return SACB_OK;
}

static int per_update_impl(sacb_handle h, int agent, const int64_t *idx, const float *prio, int64_t B, int is_final);
extern "C" int sacb_per_update(sacb_handle h, int agent, const int64_t *idx, const float *prio, int64_t B) { return per_update_impl(h, agent, idx, prio, B, 0); }
extern "C" int sacb_per_update_final(sacb_handle h, int agent, const int64_t *idx, const float *prio, int64_t B) { return per_update_impl(h, agent, idx, prio, B, 1); }
extern "C" int sacb_per_update_from_td(sacb_handle h, int agent, int64_t B) {
    if (!h || agent != 0 || h->cfg.replay_kind != SACB_REPLAY_PER) return fail(SACB_ERR_ARG, "bad argument");
    if (B < 1 || B > h->cfg.max_batch) return fail(SACB_ERR_ARG, "batch size out of range");
    SACB_CUDA(launch_pdl(per_update_kernel, dim3(1), dim3(1024), sizeof(int64_t) * B, h->stream, h->use_pdl != 0, h->prio, h->p_alpha,
                         (const int64_t *)h->last_idx_dev, (const float *)(h->ws + h->L.td), (int)B, h->cfg.per_alpha, 0));
    h->kernel_launches++;
    return SACB_OK;
}
static int per_update_impl(sacb_handle h, int agent, const int64_t *idx, const float *prio, int64_t B, int is_final) {
    if (!h || agent != 0 || h->cfg.replay_kind != SACB_REPLAY_PER || !prio) return fail(SACB_ERR_ARG, "bad argument");
    if (B < 1 || B > h->cfg.max_batch) return fail(SACB_ERR_ARG, "batch size out of range");
    PerWs w = per_ws_of(h, 0);
    const int64_t *idx_dev = h->last_idx_dev;
    if (idx) {
        for (int64_t j = 0; j < B; j++) if (idx[j] < 0 || idx[j] >= h->cfg.capacity) return fail(SACB_ERR_STATE, "priority index out of range");
        SACB_CUDA(cudaMemcpyAsync(w.idx, idx, sizeof(int64_t) * B, cudaMemcpyHostToDevice, h->stream));
        idx_dev = w.idx;
    }
    SACB_CUDA(cudaMemcpyAsync(w.weights, prio, sizeof(float) * B, cudaMemcpyHostToDevice, h->stream));
    per_update_kernel<<<1, 1024, sizeof(int64_t) * B, h->stream>>>(h->prio, h->p_alpha, idx_dev, w.weights, (int)B, h->cfg.per_alpha, is_final);
    h->kernel_launches++;
    SACB_CUDA(cudaGetLastError());
    SACB_CUDA(cudaStreamSynchronize(h->stream));
    return SACB_OK;
}

extern "C" int sacb_per_get_priorities(sacb_handle h, int agent, float *prio, int64_t n) {
    if (!h || agent != 0 || !h->prio || n > h->cfg.capacity) return fail(SACB_ERR_ARG, "bad argument");
    SACB_CUDA(cudaMemcpyAsync(prio, h->prio, sizeof(float) * n, cudaMemcpyDeviceToHost, h->stream));
    SACB_CUDA(cudaStreamSynchronize(h->stream));
    return SACB_OK;
}

extern "C" int sacb_per_set_priorities(sacb_handle h, int agent, const float *prio, const float *p_alpha, int64_t n) {
    if (!h || agent != 0 || !h->prio || n > h->cfg.capacity || !prio) return fail(SACB_ERR_ARG, "bad argument");
    SACB_CUDA(cudaMemcpyAsync(h->prio, prio, sizeof(float) * n, cudaMemcpyHostToDevice, h->stream));
    if (p_alpha) SACB_CUDA(cudaMemcpyAsync(h->p_alpha, p_alpha, sizeof(float) * n, cudaMemcpyHostToDevice, h->stream));
    else { pow_alpha_kernel<<<256, 256, 0, h->stream>>>(h->prio, h->p_alpha, n, h->cfg.per_alpha); h->kernel_launches++; }
    SACB_CUDA(cudaStreamSynchronize(h->stream));
    return SACB_OK;
}

extern "C" int sacb_per_get_stats(sacb_handle h, int agent, sacb_per_stats *out) {
    if (!h || agent != 0 || !h->prio || !out) return fail(SACB_ERR_ARG, "bad argument");
    PerWs w = per_ws_of(h, 0);
    int counters[4] = {0, 0, 0, 0};
    float tot = 0.f; double last = 0.0;
    const int64_t n = std::max<int64_t>(1, h->r_len[0]);
    const int n_chunks = (int)((n + kChunk - 1) / kChunk);
    SACB_CUDA(cudaMemcpyAsync(counters, w.counters, sizeof(int) * 3, cudaMemcpyDeviceToHost, h->stream));
    SACB_CUDA(cudaMemcpyAsync(&tot, w.total, sizeof(float), cudaMemcpyDeviceToHost, h->stream));
    SACB_CUDA(cudaMemcpyAsync(&last, w.chunk_carry + n_chunks, sizeof(double), cudaMemcpyDeviceToHost, h->stream));
    SACB_CUDA(cudaStreamSynchronize(h->stream));
    out->frame = h->per_frame[0]; out->pos = h->r_pos[0]; out->len = h->r_len[0];
    out->n_fine = counters[0]; out->n_flagged = counters[1]; out->n_exact_fallbacks = counters[2];
    out->total_f32 = tot; out->cdf_last = last;
    return SACB_OK;
}

extern "C" int sacb_per_set_frame(sacb_handle h, int agent, int64_t frame) {
    if (!h || agent != 0) return fail(SACB_ERR_ARG, "bad argument");
    h->per_frame[0] = frame;
    return SACB_OK;
}
