// GEMM tile workers of the SAC update program.
//   gemm_tile_ffma : fp32 FFMA tile (strict-parity mode, and the checker of the tensor-core tile)
//   gemm_tile_tc   : tcgen05.mma kind::tf32 tile, accumulator in TMEM, operands staged in shared memory in
//                    the canonical K-major SWIZZLE_128B layout (fp32 rows of 128 B)
// Both compute  C[M,N] = epilogue( A[M,K] * B[N,K]^T )  for one output tile of a Task.
#pragma once
#include "common.cuh"

namespace sacb {

// ---- resolved (per agent) views ---------------------------------------------------------------------------
struct OperandR {
    const float *p;
    int ld, mn_major, xform;
    const float *rvec, *cvec;
};

struct EpiR {
    int epi, M, N;
    float *C; int ldc; int accumulate;
    const float *bias;
    const float *mask; int ld_mask;
    float *w, *m, *v, *wt, *gexp;
    int apply;
    float step_size, bc2_sqrt, tau;
};

__device__ __forceinline__ float ldcg(const float *p) { return __ldcg(p); }

__device__ __forceinline__ OperandR resolve_operand(const Operand &o, const AgentBases &b, int agent) {
    OperandR r;
    r.p = resolve(o.ptr, b, agent);
    r.ld = o.ld; r.mn_major = o.mn_major; r.xform = o.xform;
    r.rvec = resolve(o.rvec, b, agent);
    r.cvec = resolve(o.cvec, b, agent);
    return r;
}

// value of the logical operand at (r, k); caller guarantees in-bounds
__device__ __forceinline__ float operand_at(const OperandR &o, int r, int k) {
    const int srow = o.mn_major ? k : r, scol = o.mn_major ? r : k;
    float v = ldcg(o.p + (int64_t)srow * o.ld + scol);
    if (o.xform) v = v > 0.f ? ldcg(o.rvec + srow) * ldcg(o.cvec + scol) : 0.f;
    return v;
}

// Adam bias corrections for step t (torch/optim/adam.py::_single_tensor_adam): python floats = double
__device__ __forceinline__ void adam_factors(int step_before, float lr, float &step_size, float &bc2_sqrt) {
    const double t = (double)(step_before + 1);
    const double bc1 = 1.0 - pow((double)0.9, t);
    const double bc2 = 1.0 - pow((double)0.999, t);
    step_size = (float)((double)lr / bc1);
    bc2_sqrt = (float)sqrt(bc2);
}

// one Adam element update (+ Polyak, sac_imp.py:146-152); returns nothing, all state in global memory
__device__ __forceinline__ void adam_element(float g, float *w, float *m, float *v, float *wt, float *gexp, int apply,
                                             float step_size, float bc2_sqrt, float tau) {
    if (gexp) *gexp = g;
    if (!apply) return;
    float mm = __ldcg(m), vv = __ldcg(v), ww = __ldcg(w);
    mm = mm + (1.0f - kBeta1) * (g - mm);                 // exp_avg.lerp_(grad, 1 - beta1)
    vv = vv * kBeta2 + (1.0f - kBeta2) * g * g;          // exp_avg_sq.mul_(beta2).addcmul_(grad, grad, 1 - beta2)
    const float denom = sqrtf(vv) / bc2_sqrt + kAdamEps;
    ww = ww - step_size * (mm / denom);
    *m = mm; *v = vv; *w = ww;
    if (wt) *wt = __ldcg(wt) * (1.0f - tau) + ww * tau;  // target <- target*(1-tau) + param*tau
}

__device__ __forceinline__ void epilogue_element(const EpiR &e, int m, int n, float acc) {
    if (m >= e.M || n >= e.N) return;
    switch (e.epi) {
        case EPI_STORE: {
            float *c = e.C + (int64_t)m * e.ldc + n;
            *c = e.accumulate ? __ldcg(c) + acc : acc;
        } break;
        case EPI_BIAS: e.C[(int64_t)m * e.ldc + n] = acc + ldcg(e.bias + n); break;
        case EPI_BIAS_RELU: e.C[(int64_t)m * e.ldc + n] = fmaxf(acc + ldcg(e.bias + n), 0.f); break;
        case EPI_MASK: e.C[(int64_t)m * e.ldc + n] = ldcg(e.mask + (int64_t)m * e.ld_mask + n) > 0.f ? acc : 0.f; break;
        case EPI_ADAM: {
            const int64_t o = (int64_t)m * e.N + n;
            adam_element(acc, e.w + o, e.m + o, e.v + o, e.wt ? e.wt + o : nullptr, e.gexp ? e.gexp + o : nullptr,
                         e.apply, e.step_size, e.bc2_sqrt, e.tau);
        } break;
    }
}

__device__ __forceinline__ EpiR resolve_epilogue(const Task &t, const AgentBases &b, int agent, const float *scalars) {
    EpiR e;
    e.epi = t.epi; e.M = t.M; e.N = t.N;
    e.C = resolve(t.C, b, agent); e.ldc = t.ldc; e.accumulate = t.accumulate;
    e.bias = resolve(t.bias, b, agent);
    e.mask = resolve(t.mask, b, agent); e.ld_mask = t.ld_mask;
    e.w = e.m = e.v = e.wt = e.gexp = nullptr; e.apply = 0; e.step_size = e.bc2_sqrt = 0.f; e.tau = 0.f;
    if (t.epi == EPI_ADAM) {
        e.w = resolve(t.adam.w, b, agent); e.m = resolve(t.adam.m, b, agent); e.v = resolve(t.adam.v, b, agent);
        e.wt = resolve(t.adam.wt, b, agent); e.gexp = resolve(t.adam.gexp, b, agent);
        e.apply = t.adam.apply; e.tau = t.adam.tau;
        adam_factors(__float_as_int(ldcg(scalars + t.adam.step_slot)), t.adam.lr, e.step_size, e.bc2_sqrt);
    }
    return e;
}

// ============================================================================================================
// fp32 FFMA tile: 64 x 64 x 16, 256 threads, 4x4 outputs per thread
// ============================================================================================================
__device__ __forceinline__ void gemm_tile_ffma(const Task &t, int tile, const AgentBases &bases, int agent,
                                               const float *scalars, float *smem) {
    const OperandR A = resolve_operand(t.A, bases, agent), B = resolve_operand(t.B, bases, agent);
    const EpiR epi = resolve_epilogue(t, bases, agent, scalars);
    const int tm = tile / t.tiles_n, tn = tile % t.tiles_n;
    const int m0 = tm * kSM, n0 = tn * kSN;
    float(*As)[kSM + 4] = reinterpret_cast<float(*)[kSM + 4]>(smem);
    float(*Bs)[kSN + 4] = reinterpret_cast<float(*)[kSN + 4]>(smem + kSK * (kSM + 4));
    const int tid = threadIdx.x, ty = (tid >> 4) & 15, tx = tid & 15;
    const bool worker = tid < 256;       // 16 x 16 threads x (4 x 4) outputs; the other warps only help with the loads
    float acc[4][4];
#pragma unroll
    for (int i = 0; i < 4; i++)
#pragma unroll
        for (int j = 0; j < 4; j++) acc[i][j] = 0.f;

    for (int k0 = 0; k0 < t.K; k0 += kSK) {
#pragma unroll
        for (int e = 0; e < kSM * kSK / kThreads; e++) {
            const int idx = tid + e * kThreads;
            int r, k;
            if (!A.mn_major) { r = idx >> 4; k = idx & 15; } else { k = idx >> 6; r = idx & 63; }
            As[k][r] = (m0 + r < t.M && k0 + k < t.K) ? operand_at(A, m0 + r, k0 + k) : 0.f;
            if (!B.mn_major) { r = idx >> 4; k = idx & 15; } else { k = idx >> 6; r = idx & 63; }
            Bs[k][r] = (n0 + r < t.N && k0 + k < t.K) ? operand_at(B, n0 + r, k0 + k) : 0.f;
        }
        __syncthreads();
#pragma unroll
        for (int k = 0; k < kSK && worker; k++) {
            const float4 a = *reinterpret_cast<const float4 *>(&As[k][ty * 4]);
            const float4 b = *reinterpret_cast<const float4 *>(&Bs[k][tx * 4]);
            const float av[4] = {a.x, a.y, a.z, a.w}, bv[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
            for (int i = 0; i < 4; i++)
#pragma unroll
                for (int j = 0; j < 4; j++) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
        }
        __syncthreads();
    }
    if (worker) {
#pragma unroll
        for (int i = 0; i < 4; i++)
#pragma unroll
            for (int j = 0; j < 4; j++) epilogue_element(epi, m0 + ty * 4 + i, n0 + tx * 4 + j, acc[i][j]);
    }
}

// ============================================================================================================
// tcgen05 tile
// ============================================================================================================
namespace tc {

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void fence_barrier_init() {
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

// bounded wait (~0.5 s): a lost arrival must surface as an error, never as a hung GPU box
__device__ __forceinline__ bool mbar_wait(uint64_t *bar, uint32_t parity, int *error_flag) {
    const uint32_t addr = smem_u32(bar);
#pragma unroll 1
    for (uint32_t it = 0; it < (1u << 23); ++it) {
        uint32_t done;
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(done) : "r"(addr), "r"(parity) : "memory");
        if (done) return true;
        if (it > 64) __nanosleep(32);
    }
    if (error_flag) atomicExch(error_flag, 1);
    return false;
}
// bounded spin on a monotonically increasing shared-memory counter (written with st.release by the MMA warp)
__device__ __forceinline__ bool counter_wait(const uint32_t *ctr, uint32_t target, int *error_flag) {
    const uint32_t addr = smem_u32(ctr);
#pragma unroll 1
    for (uint32_t it = 0; it < (1u << 23); ++it) {
        uint32_t v;
        asm volatile("ld.acquire.cta.shared::cta.u32 %0, [%1];" : "=r"(v) : "r"(addr) : "memory");
        if ((int32_t)(v - target) >= 0) return true;
        if (it > 64) __nanosleep(32);
    }
    if (error_flag) atomicExch(error_flag, 1);
    return false;
}
__device__ __forceinline__ void counter_publish(uint32_t *ctr, uint32_t v) {
    asm volatile("st.release.cta.shared::cta.u32 [%0], %1;" ::"r"(smem_u32(ctr)), "r"(v) : "memory");
}

__device__ __forceinline__ bool elect_one() {      // one lane of a converged warp
    uint32_t pred;
    asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(pred));
    return pred != 0;
}
__device__ __forceinline__ void tmem_alloc(uint32_t *dst_smem, uint32_t ncols) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(dst_smem)), "r"(ncols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t *bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
// D[tmem] (+)= A[smem] * B[smem]^T, tf32 inputs (fp32 bit patterns), fp32 accumulate
__device__ __forceinline__ void umma_tf32(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, float (&v)[16]) {
    uint32_t r[16];
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
          "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(taddr) : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
    for (int i = 0; i < 16; i++) v[i] = __uint_as_float(r[i]);
}

// UMMA shared-memory descriptor, K-major, SWIZZLE_128B (cute/arch/mma_sm100_desc.hpp::SmemDescriptor):
// start>>4 [0,14) | LBO>>4 [16,30) (ignored for swizzled K-major, 1) | SBO>>4 [32,46) = 1024 B between 8-row
// groups | version=1 [46,48) | layout_type=2 (SWIZZLE_128B) [61,64)
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr) {
    return (uint64_t)((saddr >> 4) & 0x3FFF) | (1ull << 16) | (64ull << 32) | (1ull << 46) | (2ull << 61);
}
// instruction descriptor (UMMA::InstrDescriptor): c_format F32=1 [4,6) | a_format TF32=2 [7,10) | b_format [10,13)
// | a_major=b_major=K (0) | N>>3 [17,23) | M>>4 [24,29)
__host__ __device__ constexpr uint32_t make_idesc(int M, int N) {
    return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}

// byte offset of the 16-byte chunk c (4 consecutive k) of row r inside one K-major SWIZZLE_128B operand tile
__host__ __device__ __forceinline__ uint32_t sw128_chunk_off(int r, int c) {
    return (uint32_t)((r >> 3) * 1024 + (r & 7) * 128 + ((c ^ (r & 7)) << 4));
}

// round-to-nearest fp32 -> tf32 (the MMA itself would truncate the low 13 mantissa bits: biased)
__device__ __forceinline__ float to_tf32(float x) {
    uint32_t r;
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x));
    return __uint_as_float(r);
}
__device__ __forceinline__ float4 to_tf32(float4 v) { return make_float4(to_tf32(v.x), to_tf32(v.y), to_tf32(v.z), to_tf32(v.w)); }

constexpr int kXkMax = 2048;     // longest K a rank-1 transformed operand may have (hidden_dim / batch)

// per-CTA state that survives across tiles (persistent kernel): pipeline position and TMEM base
struct TcState {
    uint32_t tmem_base;
    uint32_t g;            // k-blocks issued so far (stage = g % kTStages)
    uint32_t accum_uses;   // completed tiles (parity of the accumulator barrier)
    uint8_t *tiles;        // 1024-aligned operand ring: kTStages x kSplit x (A 16 KB | B 8 KB)
    float *xr, *xk;        // rank-1 transform vectors of the current tile: by tile row [kTM], by k [kXkMax]
    uint64_t *full_bar;    // [kTStages]  loader warp -> MMA warp: the k-block is in shared memory
    uint64_t *empty_bar;   // [kTStages]  tcgen05.commit -> loader warps: the MMAs that read the slot are done
    uint64_t *accum_bar;   //             tcgen05.commit -> everyone: the accumulator tile is complete
    uint32_t *consumed;    // k-blocks whose MMAs have completed (monotonic; published by the MMA warp, polled by loaders)
    unsigned long long *trace;   // optional per-CTA timestamps (profiling aid)
};

// how a lane fetches a 16-byte chunk (4 consecutive k of one operand row) from global memory
enum FillMode { FILL_KVEC = 0,   // K contiguous, 16 B aligned rows: one LDG.128
                FILL_KSCALAR,    // K contiguous, unaligned rows (e.g. ld = 365): 4 x LDG.32
                FILL_MN };       // MN contiguous ([K, MN] storage): 4 x LDG.32 at stride ld, coalesced across the warp

constexpr int kWarps = kThreads / 32;
constexpr int kLoaders = kWarps - 1;                       // loader warps; the last warp issues the MMAs
constexpr int kChunksA = kTM * (kTK / 4);                  // 1024 16-byte chunks of A per k-block
constexpr int kChunksAB = (kTM + kTN) * (kTK / 4);         // 1536 with B
constexpr int kPassChunks = 12;                            // chunks a lane keeps in flight per pass
constexpr int kPasses = kChunksAB / (32 * kPassChunks);    // 4
static_assert(kPasses * 32 * kPassChunks == kChunksAB, "pass geometry");

// chunk idx -> (row r, 16-byte chunk c) of the operand tile; the mapping keeps global loads coalesced:
// K-major: 8 consecutive lanes read one row's 128 B; MN-major: 32 consecutive lanes read 32 consecutive rows
template <int MODE, int ROWS>
__device__ __forceinline__ void chunk_rc(int idx, int &r, int &c) {
    if (MODE == FILL_MN) { c = idx / ROWS; r = idx % ROWS; } else { r = idx >> 3; c = idx & 7; }
}

// kv = how many of the chunk's 4 values lie inside [0,K) (>= 4: all).  Pure loads.  A K-major vector chunk is read
// whole (rows are 16 B aligned and ld >= roundup4(K)); its lanes beyond K are cleared when it is stored.
template <int MODE>
__device__ __forceinline__ float4 load_chunk(const float *src, int64_t ld, int kv) {
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (kv <= 0) return v;
    if (MODE == FILL_KVEC) return __ldcg(reinterpret_cast<const float4 *>(src));
    const int64_t st = (MODE == FILL_MN) ? ld : 1;
    v.x = ldcg(src);
    if (kv > 1) v.y = ldcg(src + st);
    if (kv > 2) v.z = ldcg(src + 2 * st);
    if (kv > 3) v.w = ldcg(src + 3 * st);
    return v;
}

__device__ __forceinline__ void sts128(uint32_t saddr, float4 v) {
    asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(saddr), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}
__device__ __forceinline__ float4 lds128(uint32_t saddr) {
    float4 v;
    asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(saddr));
    return v;
}
__device__ __forceinline__ float lds32(uint32_t saddr) {
    float v;
    asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(saddr));
    return v;
}
__device__ __forceinline__ void mbar_arrive(uint64_t *bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
// error-compensated "3xTF32": hi = x with the 13 low mantissa bits cleared (exactly what the tensor core would read),
// lo = x - hi (exact in fp32, |lo| < 2^-10 |x|; the tensor core reads its top 11 bits) => x = hi + lo up to 2^-20 |x|.
// Single-pass mode stores x as is (the MMA ignores the low 13 bits).
__device__ __forceinline__ float tf32_hi(float x) { return __uint_as_float(__float_as_uint(x) & 0xFFFFE000u); }
template <int kSplit>
__device__ __forceinline__ void store_chunk(uint32_t saddr, float4 v) {
    if (kSplit == 2) {
        const float4 hi = make_float4(tf32_hi(v.x), tf32_hi(v.y), tf32_hi(v.z), tf32_hi(v.w));
        sts128(saddr, hi);
        sts128(saddr + kTcStageBytes, make_float4(v.x - hi.x, v.y - hi.y, v.z - hi.z, v.w - hi.w));
    } else {
        sts128(saddr, v);
    }
}

// ---- main loop of one output tile -------------------------------------------------------------------------------
// Latency-parallel producer/consumer pipeline without any CTA-wide barrier:
//   * each of the kLoaders loader warps owns whole k-blocks (kb = w, w + kLoaders, ...): global -> registers (12 x 16 B
//     per lane in flight, 4 passes) -> [rank-1 transform, tf32 hi/lo split] -> swizzled smem stage -> fence -> arrive(full)
//     so up to kLoaders k-blocks of global loads are in flight per SM and nobody waits on anybody else's loads
//   * the MMA warp waits full[stage], one elected lane issues the tcgen05.mma's, tcgen05.commit -> empty[stage]
//   * mbarrier phases only carry one parity bit, and loaders run many k-blocks ahead, so loaders never wait on empty[]
//     themselves: the MMA warp (the only, in-order waiter of empty[]) publishes a monotonic `consumed` counter instead
// stage layout: [A_hi 16K | B_hi 8K] (+ [A_lo | B_lo] when kSplit == 2).  XF: operand A carries the rank-1 transform.
template <int kSplit, int AM, int BM, bool XF>
__device__ __forceinline__ void tc_mainloop(const OperandR &A, const OperandR &B, int m0, int n0, int M, int N, int K,
                                            TcState &st, int *error_flag) {
    const int tid = threadIdx.x, lane = tid & 31;
    constexpr uint32_t idesc = make_idesc(kTM, kTN);
    constexpr int kStageBytes = kSplit * kTcStageBytes;
    const int nkb = cdiv(K, kTK);
    const uint32_t tiles = smem_u32(st.tiles), xk_s = smem_u32(st.xk), xr_s = smem_u32(st.xr);
    const int warp_u = __shfl_sync(0xffffffffu, tid >> 5, 0);     // warp-uniform role index
    const uint32_t g0 = st.g;

    if (XF) {   // rank-1 operand: vector indexed by the tile row and vector indexed by k, staged once per tile
        const float *by_row = A.mn_major ? A.cvec : A.rvec, *by_k = A.mn_major ? A.rvec : A.cvec;
        for (int i = tid; i < kTM; i += kThreads) st.xr[i] = (m0 + i < M) ? ldcg(by_row + m0 + i) : 0.f;
        for (int i = tid; i < nkb * kTK; i += kThreads) st.xk[i] = (i < K) ? ldcg(by_k + i) : 0.f;
        __syncthreads();
    }

    if (warp_u < kLoaders) {
        for (int kb = warp_u; kb < nkb; kb += kLoaders) {
            const uint32_t g = g0 + kb, s = g % kTStages;
            const uint32_t stage = tiles + s * kStageBytes;
            const int k0 = kb * kTK, krem = K - k0;          // krem >= 32: full k-block
#pragma unroll 1
            for (int pass = 0; pass < kPasses; pass++) {
                float4 v[kPassChunks];
                // ---- issue all loads of the pass -----------------------------------------------------------------
#pragma unroll
                for (int j = 0; j < kPassChunks; j++) {
                    const int i = (pass * kPassChunks + j) * 32 + lane;
                    int r, c;
                    if (i < kChunksA) {
                        chunk_rc<AM, kTM>(i, r, c);
                        const int row = m0 + r;
                        const float *src = (AM == FILL_MN) ? A.p + (int64_t)(k0 + 4 * c) * A.ld + row : A.p + (int64_t)row * A.ld + k0 + 4 * c;
                        v[j] = load_chunk<AM>(src, A.ld, row < M ? krem - 4 * c : 0);
                    } else {
                        chunk_rc<BM, kTN>(i - kChunksA, r, c);
                        const int row = n0 + r;
                        const float *src = (BM == FILL_MN) ? B.p + (int64_t)(k0 + 4 * c) * B.ld + row : B.p + (int64_t)row * B.ld + k0 + 4 * c;
                        v[j] = load_chunk<BM>(src, B.ld, row < N ? krem - 4 * c : 0);
                    }
                }
                // the MMAs that last read this stage must be done before it is overwritten (loads are already in flight)
                if (pass == 0 && g >= kTStages) counter_wait(st.consumed, g - kTStages + 1, error_flag);
                // ---- transform + store ---------------------------------------------------------------------------
#pragma unroll
                for (int j = 0; j < kPassChunks; j++) {
                    const int i = (pass * kPassChunks + j) * 32 + lane;
                    int r, c;
                    float4 x = v[j];
                    if (i < kChunksA) {
                        chunk_rc<AM, kTM>(i, r, c);
                        if (AM == FILL_KVEC && krem < kTK) {     // lanes of a vector chunk beyond K
                            const int kv = krem - 4 * c;
                            if (kv < 4) { if (kv < 1) x.x = 0.f; if (kv < 2) x.y = 0.f; if (kv < 3) x.z = 0.f; x.w = 0.f; }
                        }
                        if (XF) {      // dq[b] * w_out[n] * relu'(h[b,n])
                            const float xr = lds32(xr_s + 4 * r);
                            const float4 xk = lds128(xk_s + 4 * (k0 + 4 * c));
                            x.x = x.x > 0.f ? xr * xk.x : 0.f; x.y = x.y > 0.f ? xr * xk.y : 0.f;
                            x.z = x.z > 0.f ? xr * xk.z : 0.f; x.w = x.w > 0.f ? xr * xk.w : 0.f;
                        }
                        store_chunk<kSplit>(stage + sw128_chunk_off(r, c), x);
                    } else {
                        chunk_rc<BM, kTN>(i - kChunksA, r, c);
                        if (BM == FILL_KVEC && krem < kTK) {
                            const int kv = krem - 4 * c;
                            if (kv < 4) { if (kv < 1) x.x = 0.f; if (kv < 2) x.y = 0.f; if (kv < 3) x.z = 0.f; x.w = 0.f; }
                        }
                        store_chunk<kSplit>(stage + kTM * kTK * 4 + sw128_chunk_off(r, c), x);
                    }
                }
            }
            fence_proxy_async();          // this lane's generic-proxy smem writes -> visible to the tensor-core (async) proxy
            __syncwarp();
            if (lane == 0) mbar_arrive(&st.full_bar[s]);
        }
    } else {
        constexpr int kLag = 2;      // retire two k-blocks behind the issue point: the tensor pipe always has work queued
        for (int kb = 0; kb < nkb + kLag; kb++) {
            if (kb >= kLag) {    // retire k-block kb-kLag: its MMAs are done -> its stage may be refilled
                const uint32_t gp = g0 + kb - kLag;
                mbar_wait(&st.empty_bar[gp % kTStages], (gp / kTStages) & 1, error_flag);
                if (lane == 0) counter_publish(st.consumed, gp + 1);
            }
            if (kb >= nkb) continue;
            const uint32_t g = g0 + kb, s = g % kTStages;
            mbar_wait(&st.full_bar[s], (g / kTStages) & 1, error_flag);
            if (elect_one()) {
                tc_fence_after();
                const uint64_t d0 = make_desc(tiles + s * kStageBytes);        // A_hi of this stage; the others are constant offsets
                constexpr uint64_t kB = (kTM * kTK * 4) >> 4, kLo = kTcStageBytes >> 4;
#pragma unroll
                for (int kk = 0; kk < kTK / 8; kk++) {  // UMMA_K = 8 tf32 = 32 B: advance the start address inside the swizzled row
                    const uint64_t da = d0 + 2 * kk, db = d0 + kB + 2 * kk;
                    if (kSplit == 2) {   // small terms first: a_lo*b_hi + a_hi*b_lo, then a_hi*b_hi
                        umma_tf32(st.tmem_base, da + kLo, db, idesc, (kb | kk) ? 1u : 0u);
                        umma_tf32(st.tmem_base, da, db + kLo, idesc, 1u);
                        umma_tf32(st.tmem_base, da, db, idesc, 1u);
                    } else {
                        umma_tf32(st.tmem_base, da, db, idesc, (kb | kk) ? 1u : 0u);
                    }
                }
                umma_commit(&st.empty_bar[s]);
                if (kb == nkb - 1) umma_commit(st.accum_bar);
            }
            __syncwarp();
        }
    }
    st.g += nkb;
}

// ---- epilogue of one thread: 16 consecutive columns of one output row; loads batched ahead of any store ----
template <int EPI>
__device__ __forceinline__ void epilogue16(const EpiR &e, int m, int n, const float (&acc)[16]) {
    if (m >= e.M || n >= e.N) return;
    const bool full = n + 15 < e.N;
    if (EPI == EPI_ADAM) {
#pragma unroll 1
        for (int h8 = 0; h8 < 16; h8 += 8) {        // two batches of 8 columns: loads of a batch are issued before its stores
            const int64_t o = (int64_t)m * e.N + n + h8;
            float w[8], mm[8], vv[8], wt[8];
#pragma unroll
            for (int j = 0; j < 8; j++) {
                const bool ok = (full || n + h8 + j < e.N) && e.apply;
                w[j] = ok ? __ldcg(e.w + o + j) : 0.f;
                mm[j] = ok ? __ldcg(e.m + o + j) : 0.f;
                vv[j] = ok ? __ldcg(e.v + o + j) : 0.f;
                wt[j] = (ok && e.wt) ? __ldcg(e.wt + o + j) : 0.f;
            }
#pragma unroll
            for (int j = 0; j < 8; j++) {
                if (!(full || n + h8 + j < e.N)) continue;
                const float g = acc[h8 + j];
                if (e.gexp) e.gexp[o + j] = g;
                if (!e.apply) continue;
                const float m1 = mm[j] + (1.0f - kBeta1) * (g - mm[j]);
                const float v1 = vv[j] * kBeta2 + (1.0f - kBeta2) * g * g;
                const float w1 = w[j] - e.step_size * (m1 / (sqrtf(v1) / e.bc2_sqrt + kAdamEps));
                e.m[o + j] = m1; e.v[o + j] = v1; e.w[o + j] = w1;
                if (e.wt) e.wt[o + j] = wt[j] * (1.0f - e.tau) + w1 * e.tau;
            }
        }
        return;
    }
    float aux[16];
    if (EPI == EPI_BIAS || EPI == EPI_BIAS_RELU) {
#pragma unroll
        for (int j = 0; j < 16; j++) aux[j] = (full || n + j < e.N) ? ldcg(e.bias + n + j) : 0.f;
    } else if (EPI == EPI_MASK) {
#pragma unroll
        for (int j = 0; j < 16; j++) aux[j] = (full || n + j < e.N) ? ldcg(e.mask + (int64_t)m * e.ld_mask + n + j) : 0.f;
    } else if (EPI == EPI_STORE) {
#pragma unroll
        for (int j = 0; j < 16; j++) aux[j] = (e.accumulate && (full || n + j < e.N)) ? __ldcg(e.C + (int64_t)m * e.ldc + n + j) : 0.f;
    }
    float out[16];
#pragma unroll
    for (int j = 0; j < 16; j++) {
        if (EPI == EPI_BIAS) out[j] = acc[j] + aux[j];
        else if (EPI == EPI_BIAS_RELU) out[j] = fmaxf(acc[j] + aux[j], 0.f);
        else if (EPI == EPI_MASK) out[j] = aux[j] > 0.f ? acc[j] : 0.f;
        else out[j] = acc[j] + aux[j];
    }
    float *c = e.C + (int64_t)m * e.ldc + n;
    if (full && (e.ldc % 4 == 0) && ((reinterpret_cast<uintptr_t>(c) & 15) == 0)) {
#pragma unroll
        for (int j = 0; j < 16; j += 4) *reinterpret_cast<float4 *>(c + j) = make_float4(out[j], out[j + 1], out[j + 2], out[j + 3]);
    } else {
#pragma unroll
        for (int j = 0; j < 16; j++) if (full || n + j < e.N) c[j] = out[j];
    }
}

template <int EPI>
__device__ __forceinline__ void tc_epilogue(const EpiR &epi, int m0, int n0, TcState &st, int *error_flag) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    mbar_wait(st.accum_bar, st.accum_uses & 1, error_flag);
    st.accum_uses++;
    tc_fence_after();
    const int row = (warp & 3) * 32 + lane;          // warp w owns TMEM lanes 32*(w%4).., column group w/4
    constexpr int kColsPerWarp = kTN / (kThreads / 128);
    const int colh = (warp >> 2) * kColsPerWarp;
#pragma unroll 1
    for (int cc = 0; cc < kColsPerWarp; cc += 16) {
        float v[16];
        tmem_ld16(st.tmem_base + ((uint32_t)((warp & 3) * 32) << 16) + (uint32_t)(colh + cc), v);
        epilogue16<EPI>(epi, m0 + row, n0 + colh + cc, v);
    }
    tc_fence_before();
    __syncthreads();     // all TMEM reads retired before the next tile's first MMA overwrites the accumulator
}

}  // namespace tc

template <int kSplit>
__device__ __forceinline__ void gemm_tile_tc(const Task &t, int tile, const AgentBases &bases, int agent,
                                             const float *scalars, tc::TcState &st, int *error_flag) {
    using namespace tc;
    const int tm = tile / t.tiles_n, tn = tile % t.tiles_n;
    const int m0 = tm * kTM, n0 = tn * kTN;
    {
        const OperandR A = resolve_operand(t.A, bases, agent), B = resolve_operand(t.B, bases, agent);
        const bool b_aligned = (B.ld % 4 == 0) && ((reinterpret_cast<uintptr_t>(B.p) & 15) == 0);
        // operand fetch modes that occur in the update program (the builder keeps every K-major A operand 16 B aligned;
        // weights as the K-major B operand may have unaligned rows, e.g. Q fc1 with ld = obs+act)
#define SACB_ML(AM_, BM_, XF_) tc_mainloop<kSplit, AM_, BM_, XF_>(A, B, m0, n0, t.M, t.N, t.K, st, error_flag)
        if (!A.mn_major) {
            if (B.mn_major) { if (A.xform) SACB_ML(FILL_KVEC, FILL_MN, true); else SACB_ML(FILL_KVEC, FILL_MN, false); }
            else if (b_aligned) SACB_ML(FILL_KVEC, FILL_KVEC, false);
            else SACB_ML(FILL_KVEC, FILL_KSCALAR, false);
        } else {
            if (A.xform) SACB_ML(FILL_MN, FILL_MN, true); else SACB_ML(FILL_MN, FILL_MN, false);
        }
#undef SACB_ML
    }
    auto stamp = [&](int slot) {
        if (st.trace && threadIdx.x == 0 && tile == (int)blockIdx.x - t.tile_begin) {
            unsigned long long ts;
            asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(ts));
            st.trace[(size_t)blockIdx.x * 8 + slot] = ts;
        }
    };
    stamp(2);
    const EpiR epi = resolve_epilogue(t, bases, agent, scalars);
    switch (t.epi) {
        case EPI_STORE: tc_epilogue<EPI_STORE>(epi, m0, n0, st, error_flag); break;
        case EPI_BIAS: tc_epilogue<EPI_BIAS>(epi, m0, n0, st, error_flag); break;
        case EPI_BIAS_RELU: tc_epilogue<EPI_BIAS_RELU>(epi, m0, n0, st, error_flag); break;
        case EPI_MASK: tc_epilogue<EPI_MASK>(epi, m0, n0, st, error_flag); break;
        default: tc_epilogue<EPI_ADAM>(epi, m0, n0, st, error_flag); break;
    }
    stamp(3);
}

}  // namespace sacb
