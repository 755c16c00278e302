// Row-wise network evaluation (select_action, QNetwork / GaussianPolicy callables) and the tensor-core self test.
#include <cstring>
#include <vector>
#include <algorithm>

#include "handle.h"
#include "tasks.cuh"

namespace sacb {

struct MlpArgs {
    const float *w[4], *b[4];
    const float *w_out, *b_out;
    int n_hidden, in_dim, hidden, out_dim;
};

// one CTA per input row: y = W_out . relu(... relu(W_0 x + b_0) ...) + b_out ; warp-per-output-row dot products,
// activations in shared memory (2 x hidden floats + in_dim).  B = 1 is the select_action latency path (sac_imp.py:54-72).
__global__ void __launch_bounds__(512) mlp_rows_kernel(MlpArgs a, const float *x, int ldx, const float *x2, int n2, float *out, int ldo) {
    extern __shared__ float sm[];
    float *cur = sm, *nxt = sm + max(a.in_dim, a.hidden);
    const int row = blockIdx.x, warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nw = blockDim.x >> 5;
    const int n1 = a.in_dim - n2;
    for (int j = threadIdx.x; j < a.in_dim; j += blockDim.x) cur[j] = j < n1 ? x[(int64_t)row * ldx + j] : x2[(int64_t)row * n2 + (j - n1)];
    __syncthreads();
    int in = a.in_dim;
    for (int l = 0; l < a.n_hidden; l++) {
        for (int o = warp; o < a.hidden; o += nw) {
            const float *wr = a.w[l] + (int64_t)o * in;
            float s = 0.f;
            for (int j = lane; j < in; j += 32) s = fmaf(__ldg(wr + j), cur[j], s);
            s = warp_sum(s);
            if (lane == 0) nxt[o] = fmaxf(s + __ldg(a.b[l] + o), 0.f);
        }
        __syncthreads();
        float *t = cur; cur = nxt; nxt = t;
        in = a.hidden;
    }
    for (int o = warp; o < a.out_dim; o += nw) {
        const float *wr = a.w_out + (int64_t)o * in;
        float s = 0.f;
        for (int j = lane; j < in; j += 32) s = fmaf(__ldg(wr + j), cur[j], s);
        s = warp_sum(s);
        if (lane == 0) out[(int64_t)row * ldo + o] = s + __ldg(a.b_out + o);
    }
}

// head_raw [n, 2A] -> action [n, A]: evaluate => tanh(mean)*scale+bias (sac_imp.py:61-64), else policy.sample (:70)
__global__ void action_kernel(const float *head, const float *eps, int n, int A, int evaluate, float scale, float bias, float *act,
                              uint64_t seed, uint32_t counter) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n * A) return;
    const int r = i / A, c = i % A;
    const float mean = head[(int64_t)r * 2 * A + c], ls = head[(int64_t)r * 2 * A + A + c];
    if (evaluate) { act[i] = tanhf(mean) * scale + bias; return; }
    const float e = eps ? eps[i] : philox_normal(seed, 0x5e1ec7u, counter, (uint32_t)r, (uint32_t)c);
    act[i] = sample_elem(mean, ls, e, scale, bias).action;
}

static MlpArgs mlp_args(sacb_handle h, int agent, int net) {
    const NetLayout &nl = net == SACB_NET_POLICY ? h->L.pol : h->L.q;
    const float *base = h->arena + (int64_t)agent * h->L.arena_size + h->L.param[net];
    MlpArgs a;
    memset(&a, 0, sizeof(a));
    for (int l = 0; l < nl.n_hidden; l++) { a.w[l] = base + nl.w[l]; a.b[l] = base + nl.b[l]; }
    a.w_out = base + nl.w_out; a.b_out = base + nl.b_out;
    a.n_hidden = nl.n_hidden; a.in_dim = nl.in_dim; a.hidden = nl.hidden; a.out_dim = nl.out_dim;
    return a;
}

static size_t mlp_smem(const MlpArgs &a) { return sizeof(float) * 2 * std::max(a.in_dim, a.hidden); }

}  // namespace sacb
using namespace sacb;

// scratch inside the workspace that the update does not need between steps: head_raw / g_head / eps regions are
// step-local, but a concurrent select_action must not clobber them -> use the pinned buffer + stage_rows instead.
extern "C" int sacb_select_action(sacb_handle h, int agent, const float *obs, int evaluate, const float *eps, float *action_out) {
    if (!h || !obs || !action_out || agent < 0 || agent >= h->cfg.n_agents) return fail(SACB_ERR_ARG, "bad argument");
    const int O = h->cfg.obs_dim, A = h->cfg.act_dim;
    float *d_obs = h->stage_rows, *d_head = d_obs + align_up(O, 4), *d_eps = d_head + align_up(2 * A, 4), *d_act = d_eps + align_up(A, 4);
    static_assert(sizeof(float) == 4, "");
    SACB_CUDA(cudaMemcpyAsync(d_obs, obs, sizeof(float) * O, cudaMemcpyHostToDevice, h->stream));
    if (eps && !evaluate) SACB_CUDA(cudaMemcpyAsync(d_eps, eps, sizeof(float) * A, cudaMemcpyHostToDevice, h->stream));
    MlpArgs a = mlp_args(h, agent, SACB_NET_POLICY);
    mlp_rows_kernel<<<1, 512, mlp_smem(a), h->stream>>>(a, d_obs, O, nullptr, 0, d_head, 2 * A);
    static uint32_t counter = 0;
    action_kernel<<<1, std::max(32, (A + 31) / 32 * 32), 0, h->stream>>>(d_head, (eps && !evaluate) ? d_eps : nullptr, 1, A, evaluate, h->cfg.action_scale,
                                                                        h->cfg.action_bias, d_act, h->cfg.seed, counter++);
    h->kernel_launches += 2;
    SACB_CUDA(cudaGetLastError());
    SACB_CUDA(cudaMemcpyAsync(action_out, d_act, sizeof(float) * A, cudaMemcpyDeviceToHost, h->stream));
    SACB_CUDA(cudaStreamSynchronize(h->stream));
    return SACB_OK;
}

extern "C" int sacb_q_forward(sacb_handle h, int agent, int net, const float *s, const float *a, int64_t n, float *q_out) {
    if (!h || !s || !a || !q_out || net < SACB_NET_Q1 || net > SACB_NET_Q2_TARGET || agent < 0 || agent >= h->cfg.n_agents) return fail(SACB_ERR_ARG, "bad argument");
    const int O = h->cfg.obs_dim, A = h->cfg.act_dim;
    float *d_s, *d_a, *d_q;
    SACB_CUDA(cudaMalloc(&d_s, sizeof(float) * n * O)); SACB_CUDA(cudaMalloc(&d_a, sizeof(float) * n * A)); SACB_CUDA(cudaMalloc(&d_q, sizeof(float) * n));
    SACB_CUDA(cudaMemcpyAsync(d_s, s, sizeof(float) * n * O, cudaMemcpyHostToDevice, h->stream));
    SACB_CUDA(cudaMemcpyAsync(d_a, a, sizeof(float) * n * A, cudaMemcpyHostToDevice, h->stream));
    MlpArgs m = mlp_args(h, agent, net);
    mlp_rows_kernel<<<(int)n, 512, mlp_smem(m), h->stream>>>(m, d_s, O, d_a, A, d_q, 1);
    h->kernel_launches++;
    SACB_CUDA(cudaMemcpyAsync(q_out, d_q, sizeof(float) * n, cudaMemcpyDeviceToHost, h->stream));
    SACB_CUDA(cudaStreamSynchronize(h->stream));
    cudaFree(d_s); cudaFree(d_a); cudaFree(d_q);
    return SACB_OK;
}

extern "C" int sacb_policy_forward(sacb_handle h, int agent, const float *s, int64_t n, float *mean_out, float *log_std_out) {
    if (!h || !s || !mean_out || !log_std_out || agent < 0 || agent >= h->cfg.n_agents) return fail(SACB_ERR_ARG, "bad argument");
    const int O = h->cfg.obs_dim, A = h->cfg.act_dim;
    float *d_s, *d_h;
    SACB_CUDA(cudaMalloc(&d_s, sizeof(float) * n * O)); SACB_CUDA(cudaMalloc(&d_h, sizeof(float) * n * 2 * A));
    SACB_CUDA(cudaMemcpyAsync(d_s, s, sizeof(float) * n * O, cudaMemcpyHostToDevice, h->stream));
    MlpArgs m = mlp_args(h, agent, SACB_NET_POLICY);
    mlp_rows_kernel<<<(int)n, 512, mlp_smem(m), h->stream>>>(m, d_s, O, nullptr, 0, d_h, 2 * A);
    h->kernel_launches++;
    std::vector<float> head((size_t)n * 2 * A);
    SACB_CUDA(cudaMemcpyAsync(head.data(), d_h, sizeof(float) * n * 2 * A, cudaMemcpyDeviceToHost, h->stream));
    SACB_CUDA(cudaStreamSynchronize(h->stream));
    for (int64_t r = 0; r < n; r++)
        for (int c = 0; c < A; c++) {
            mean_out[r * A + c] = head[r * 2 * A + c];
            log_std_out[r * A + c] = std::min(std::max(head[r * 2 * A + A + c], kLogStdMin), kLogStdMax);   // torch.clamp(log_std, -20, 2)
        }
    cudaFree(d_s); cudaFree(d_h);
    return SACB_OK;
}

// ---- data-parallel mode: implemented in dp.cu ---------------------------------------------------------------------

// ---- self test: tensor-core tile vs FFMA tile on random operands ------------------------------------------------
namespace sacb {
template <int kMath>
__global__ void __launch_bounds__(kThreads, 1) gemm_selftest_kernel(Task t, AgentBases bases, int *error_flag) {
    extern __shared__ __align__(16) uint8_t smem_raw[];
    __shared__ uint64_t s_bars[2 * kTStages + 1];
    __shared__ uint32_t s_tmem;
    __shared__ uint32_t s_consumed;
    tc::TcState st;
    st.g = 0; st.accum_uses = 0; st.tmem_base = 0; st.full_bar = s_bars; st.empty_bar = s_bars + kTStages; st.accum_bar = s_bars + 2 * kTStages; st.consumed = &s_consumed; st.trace = nullptr;
    constexpr int kSplit = kMath == SACB_MATH_TF32X3 ? 2 : 1;
    constexpr bool kTc = kMath != SACB_MATH_FP32;
    st.tiles = reinterpret_cast<uint8_t *>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    st.xr = reinterpret_cast<float *>(st.tiles + kTStages * kSplit * kTcStageBytes);
    st.xk = st.xr + kTM;
    if (kTc) {
        if (threadIdx.x == 0) { for (int i = 0; i <= 2 * kTStages; i++) tc::mbar_init(&s_bars[i], 1);
            s_consumed = 0; tc::fence_barrier_init(); tc::fence_proxy_async(); }
        if (threadIdx.x < 32) tc::tmem_alloc(&s_tmem, kTN);
        tc::tc_fence_before(); __syncthreads(); tc::tc_fence_after();
        st.tmem_base = s_tmem;
    }
    for (int tile = blockIdx.x; tile < t.n_tiles; tile += gridDim.x) {
        if (kTc) gemm_tile_tc<kSplit>(t, tile, bases, 0, nullptr, st, error_flag);
        else gemm_tile_ffma(t, tile, bases, 0, nullptr, reinterpret_cast<float *>(smem_raw));
    }
    if (kTc) { tc::tc_fence_before(); __syncthreads(); if (threadIdx.x < 32) tc::tmem_dealloc(st.tmem_base, kTN); }
}
}  // namespace sacb

extern "C" int sacb_selftest_gemm(int device, int M, int N, int K, int a_mn, int b_mn_and_mode, float *rel_err_out) {
    // b_mn_and_mode: bit 0 = B operand MN-major, bit 1 = use the error-compensated 3xTF32 tile instead of plain tf32
    const int b_mn = b_mn_and_mode & 1, x3 = (b_mn_and_mode >> 1) & 1;
    if (M < 1 || N < 1 || K < 1 || !rel_err_out) return fail(SACB_ERR_ARG, "bad argument");
    SACB_CUDA(cudaSetDevice(device));
    const int lda = a_mn ? (int)align_up(M, 4) + 4 : (int)align_up(K, 4) + 4, ldb = b_mn ? (int)align_up(N, 4) + 4 : (int)align_up(K, 4) + 4;
    const int64_t na = (int64_t)(a_mn ? K : M) * lda, nb = (int64_t)(b_mn ? K : N) * ldb, nc = (int64_t)M * N;
    std::vector<float> ha(na), hb(nb), c0(nc), c1(nc);
    uint32_t s = 12345u;
    auto rnd = [&]() { s = s * 1664525u + 1013904223u; return ((s >> 8) & 0xFFFF) / 65536.0f - 0.5f; };
    for (auto &v : ha) v = rnd();
    for (auto &v : hb) v = rnd();
    float *d;
    const int64_t total = na + nb + 2 * nc + 64;
    SACB_CUDA(cudaMalloc(&d, sizeof(float) * total));
    SACB_CUDA(cudaMemset(d, 0, sizeof(float) * total));
    int *flag;
    SACB_CUDA(cudaMalloc(&flag, 64)); SACB_CUDA(cudaMemset(flag, 0, 64));
    const int64_t oa = 0, ob = align_up(na, 4), oc0 = ob + align_up(nb, 4), oc1 = oc0 + align_up(nc, 4);
    SACB_CUDA(cudaMemcpy(d + oa, ha.data(), sizeof(float) * na, cudaMemcpyHostToDevice));
    SACB_CUDA(cudaMemcpy(d + ob, hb.data(), sizeof(float) * nb, cudaMemcpyHostToDevice));
    AgentBases bases{d, d, 0, 0};
    auto mk = [&](int tm, int tn, int64_t oc) {
        Task t; memset(&t, 0, sizeof(t));
        t.type = T_GEMM; t.M = M; t.N = N; t.K = K; t.epi = EPI_STORE;
        t.A.ptr = make_ref(0, oa); t.A.ld = lda; t.A.mn_major = a_mn; t.A.rvec = t.A.cvec = null_ref();
        t.B.ptr = make_ref(0, ob); t.B.ld = ldb; t.B.mn_major = b_mn; t.B.rvec = t.B.cvec = null_ref();
        t.C = make_ref(0, oc); t.ldc = N; t.bias = t.mask = null_ref();
        t.tiles_m = cdiv(M, tm); t.tiles_n = cdiv(N, tn); t.n_tiles = t.tiles_m * t.tiles_n;
        return t;
    };
    Task t0 = mk(kSM, kSN, oc0), t1 = mk(kTM, kTN, oc1);
    gemm_selftest_kernel<SACB_MATH_FP32><<<t0.n_tiles, kThreads, kSimtSmemBytes>>>(t0, bases, flag);
    if (x3) {
        SACB_CUDA(cudaFuncSetAttribute(gemm_selftest_kernel<SACB_MATH_TF32X3>, cudaFuncAttributeMaxDynamicSharedMemorySize, tc_smem_bytes(2)));
        gemm_selftest_kernel<SACB_MATH_TF32X3><<<std::min(t1.n_tiles, 148), kThreads, tc_smem_bytes(2)>>>(t1, bases, flag);
    } else {
        SACB_CUDA(cudaFuncSetAttribute(gemm_selftest_kernel<SACB_MATH_TF32>, cudaFuncAttributeMaxDynamicSharedMemorySize, tc_smem_bytes(1)));
        gemm_selftest_kernel<SACB_MATH_TF32><<<std::min(t1.n_tiles, 148), kThreads, tc_smem_bytes(1)>>>(t1, bases, flag);
    }
    SACB_CUDA(cudaDeviceSynchronize());
    int hf = 0;
    SACB_CUDA(cudaMemcpy(&hf, flag, sizeof(int), cudaMemcpyDeviceToHost));
    SACB_CUDA(cudaMemcpy(c0.data(), d + oc0, sizeof(float) * nc, cudaMemcpyDeviceToHost));
    SACB_CUDA(cudaMemcpy(c1.data(), d + oc1, sizeof(float) * nc, cudaMemcpyDeviceToHost));
    cudaFree(d); cudaFree(flag);
    // host reference in double on a sample of entries pins the FFMA tile itself
    double maxref = 0, maxdiff = 0, maxdiff_host = 0;
    for (int64_t i = 0; i < nc; i++) { maxref = std::max(maxref, (double)fabsf(c0[i])); maxdiff = std::max(maxdiff, (double)fabsf(c0[i] - c1[i])); }
    for (int64_t i = 0; i < nc; i += std::max<int64_t>(1, nc / 257)) {
        const int m = (int)(i / N), n = (int)(i % N);
        double acc = 0;
        for (int k = 0; k < K; k++) acc += (double)(a_mn ? ha[(int64_t)k * lda + m] : ha[(int64_t)m * lda + k]) * (double)(b_mn ? hb[(int64_t)k * ldb + n] : hb[(int64_t)n * ldb + k]);
        maxdiff_host = std::max(maxdiff_host, fabs(acc - (double)c0[i]));
    }
    if (hf) return fail(SACB_ERR_DEVICE, "tcgen05 pipeline watchdog fired in the self test");
    if (maxdiff_host > 1e-4 * std::max(1.0, maxref)) return fail(SACB_ERR_DEVICE, "FFMA tile disagrees with the host reference");
    *rel_err_out = (float)(maxdiff / std::max(maxref, 1e-30));
    return SACB_OK;
}
