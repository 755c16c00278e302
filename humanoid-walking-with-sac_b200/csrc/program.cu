// The SAC update as a static tile program: layout, kernel, host-side program builder, launch / CUDA graph.
//
// Reference semantics: sac_imp.SAC.update_parameters (sac_imp.py:74-144), in its order of operations
// (SURVEY 3.2): target with pre-step policy/targets -> twin critic MSE + Adam -> actor through the UPDATED
// critics + Adam -> temperature + Adam -> Polyak.
#include <tuple>
#include <cstring>
#include <cstdio>
#include <cstdlib>

#include "handle.h"
#include "tasks.cuh"

namespace sacb {

// ================================================================================================================
// layout
// ================================================================================================================
void NetLayout::tensor(int t, int64_t &off, int64_t &rows, int64_t &cols) const {
    const int nh2 = 2 * n_hidden;
    if (t < nh2) {
        const int l = t / 2;
        if (t % 2 == 0) { off = w[l]; rows = hidden; cols = in_of(l); }
        else { off = b[l]; rows = hidden; cols = 1; }
        return;
    }
    const int k = t - nh2;
    if (!is_policy) {   // fc_out.weight [1,H], fc_out.bias [1]
        if (k == 0) { off = w_out; rows = 1; cols = hidden; } else { off = b_out; rows = 1; cols = 1; }
    } else {            // mean.weight, mean.bias, log_std.weight, log_std.bias
        const int A = out_dim / 2;
        if (k == 0) { off = w_out; rows = A; cols = hidden; }
        else if (k == 1) { off = b_out; rows = A; cols = 1; }
        else if (k == 2) { off = w_out + (int64_t)A * hidden; rows = A; cols = hidden; }
        else { off = b_out + A; rows = A; cols = 1; }
    }
}

static void build_net(NetLayout &n, bool is_policy, int in_dim, int hidden, int n_hidden, int out_dim) {
    n.is_policy = is_policy; n.in_dim = in_dim; n.hidden = hidden; n.n_hidden = n_hidden; n.out_dim = out_dim;
    int64_t o = 0;
    for (int l = 0; l < n_hidden; l++) {
        n.w[l] = o; o = align_up(o + (int64_t)hidden * n.in_of(l), 4);
        n.b[l] = o; o = align_up(o + hidden, 4);
    }
    n.w_out = o; o = align_up(o + (int64_t)out_dim * hidden, 4);
    n.b_out = o; o = align_up(o + out_dim, 4);
    n.size = align_up(o, 32);
}

void Layout::build(int obs_, int act_, int hidden_, int n_hidden_, int maxB_) {
    obs = obs_; act = act_; hidden = hidden_; n_hidden = n_hidden_; maxB = maxB_;
    ldx = (int)align_up(obs + act, 4);
    build_net(pol, true, obs, hidden, n_hidden, 2 * act);
    build_net(q, false, obs + act, hidden, n_hidden, 1);
    int64_t o = 0;
    scalars = o; o += 32;
    param[0] = o; o += pol.size;
    for (int k = 1; k < 5; k++) { param[k] = o; o += q.size; }
    const int64_t sz[3] = {pol.size, q.size, q.size};
    for (int k = 0; k < 3; k++) { adam_m[k] = o; o += sz[k]; }
    for (int k = 0; k < 3; k++) { adam_v[k] = o; o += sz[k]; }
    for (int k = 0; k < 3; k++) { grad[k] = o; o += sz[k]; }
    grad_scalars = o; o += 32;
    arena_size = align_up(o, 64);
    // workspace
    const int64_t B = maxB, H = hidden, A = act;
    auto take = [&](int64_t n) { int64_t r = o; o = align_up(o + n, 32); return r; };
    o = 0;
    X = take(3 * B * ldx);
    r = take(B); d = take(B); isw = take(B); y = take(B); td = take(B);
    dq[0] = take(B); dq[1] = take(B); dqa[0] = take(B); dqa[1] = take(B);
    logp = take(2 * B); eps = take(2 * B * A);
    head_raw = take(2 * B * 2 * A); ldg = (int)align_up(2 * A, 4); g_head = take(B * ldg);
    da[0] = take(B * A); da[1] = take(B * A);
    wsnap[0] = take(H); wsnap[1] = take(H);
    loss_part = take(2 * ((B + 7) / 8) + 8); aloss_part = take(2 * ((B + 7) / 8) + 8);
    for (int l = 0; l < n_hidden; l++) { hp[l] = take(2 * B * H); dhp[l] = take(B * H); }
    for (int k = 0; k < 2; k++)
        for (int l = 0; l < n_hidden; l++) {
            ht[k][l] = take(B * H); hc[k][l] = take(B * H); ha[k][l] = take(B * H);
            dhc[k][l] = take(B * H); dha[k][l] = take(B * H);
        }
    ws_size = align_up(o, 64);
}

// ================================================================================================================
// kernel
// ================================================================================================================
__device__ __forceinline__ void grid_barrier(unsigned int *counter, unsigned int target, int *error_flag) {
    __syncthreads();
    if (threadIdx.x == 0) {
        __threadfence();
        atomicAdd(counter, 1u);
        unsigned int it = 0;
        while (true) {
            unsigned int v;
            asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(counter) : "memory");
            if (v >= target) break;
            if (++it > (1u << 24)) { atomicExch(error_flag, 2); break; }   // watchdog: never hang the box
        }
        __threadfence();
    }
    __syncthreads();
}

template <int kMath>
__global__ void __launch_bounds__(kThreads, 1)
sac_update_kernel(const Program P, const int stage_begin, const int stage_end, const int tc_setup, const uint64_t seed) {
    extern __shared__ __align__(16) uint8_t smem_raw[];
    __shared__ uint64_t s_bars[2 * kTStages + 1];
    __shared__ uint32_t s_tmem;
    __shared__ uint32_t s_consumed;
    __shared__ float s_red[kThreads];

    auto stamp = [&](int slot) {
        if (P.trace && threadIdx.x == 0) {
            unsigned long long t;
            asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
            P.trace[(size_t)blockIdx.x * 8 + slot] = t;
        }
    };
    stamp(0);
    tc::TcState st;
    st.g = 0; st.accum_uses = 0; st.tmem_base = 0;
    st.full_bar = s_bars; st.empty_bar = s_bars + kTStages; st.accum_bar = s_bars + 2 * kTStages; st.consumed = &s_consumed; st.trace = P.trace;
    constexpr int kSplit = kMath == SACB_MATH_TF32X3 ? 2 : 1;
    constexpr bool kTc = kMath != SACB_MATH_FP32;
    st.tiles = reinterpret_cast<uint8_t *>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    st.xr = reinterpret_cast<float *>(st.tiles + kTStages * kSplit * kTcStageBytes);
    st.xk = st.xr + kTM;
    if (kTc && tc_setup) {
        if (threadIdx.x == 0) {
            for (int i = 0; i <= 2 * kTStages; i++) tc::mbar_init(&s_bars[i], 1);
            s_consumed = 0;
            tc::fence_barrier_init();
            tc::fence_proxy_async();
        }
        if (threadIdx.x < 32) tc::tmem_alloc(&s_tmem, kTN);
        tc::tc_fence_before();
        __syncthreads();
        tc::tc_fence_after();
        st.tmem_base = s_tmem;
    }

    stamp(1);
    unsigned int bar_target = 0;
    for (int s = stage_begin; s < stage_end; s++) {
        const Stage sg = P.stages[s];
        const int total = sg.n_tiles * P.n_agents;
        for (int wi = blockIdx.x; wi < total; wi += gridDim.x) {
            const int agent = wi / sg.n_tiles, tile_in_stage = wi % sg.n_tiles;
            int ti = sg.task_begin;
            while (ti + 1 < sg.task_end && tile_in_stage >= P.tasks[ti + 1].tile_begin) ti++;
            const Task &t = P.tasks[ti];
            const int tile = tile_in_stage - t.tile_begin;
            float *scalars = resolve(P.scalars, P.bases, agent);
            switch (t.type) {
                case T_GEMM:
                    if (kTc) gemm_tile_tc<kSplit>(t, tile, P.bases, agent, scalars, st, P.error_flag);
                    else gemm_tile_ffma(t, tile, P.bases, agent, scalars, reinterpret_cast<float *>(smem_raw));
                    break;
                case T_GATHER: task_gather(t, tile, P, agent); break;
                case T_SAMPLE: task_sample(t, tile, P, agent, scalars, seed); break;
                case T_TARGET_LOSS: task_target_loss(t, tile, P, agent, scalars, s_red); break;
                case T_ACTOR_LOSS: task_actor_loss(t, tile, P, agent, scalars, s_red); break;
                case T_SAMPLE_BWD: task_sample_bwd(t, tile, P, agent, scalars); break;
                case T_OUT_ADAM: task_out_adam(t, tile, P, agent, scalars, s_red); break;
                case T_BIAS_ADAM: task_bias_adam(t, tile, P, agent, scalars, s_red); break;
                case T_FINISH: task_finish(t, P, agent, scalars, s_red); break;
            }
            if (wi == (int)blockIdx.x) stamp(4);
        }
        if (s + 1 < stage_end) {
            bar_target += gridDim.x;
            grid_barrier(P.barrier, bar_target, P.error_flag);
        }
    }

    if (kTc && tc_setup) {
        tc::tc_fence_before();
        __syncthreads();
        if (threadIdx.x < 32) tc::tmem_dealloc(st.tmem_base, kTN);
    }
    stamp(5);
}

// ================================================================================================================
// program builder
// ================================================================================================================
namespace {

struct Builder {
    sacb_handle h;
    const Layout &L;
    ProgramKey key;
    std::vector<Task> tasks;
    std::vector<Stage> stages;
    std::vector<int> has_gemm;
    int tm, tn;   // GEMM tile dims of the math mode
    bool stage_open = false;

    Builder(sacb_handle h_, const ProgramKey &k) : h(h_), L(h_->L), key(k) {
        if (math_is_tc(h->cfg.math_mode)) { tm = kTM; tn = kTN; } else { tm = kSM; tn = kSN; }
    }
    static Ref A(int64_t off) { return make_ref(0, off); }
    static Ref W(int64_t off) { return make_ref(1, off); }

    void begin_stage() {
        Stage s{}; s.task_begin = (int)tasks.size(); s.task_end = s.task_begin; s.n_tiles = 0;
        stages.push_back(s); has_gemm.push_back(0);
    }
    Task blank(int type) {
        Task t; memset(&t, 0, sizeof(t));
        t.type = type;
        t.C = t.bias = t.mask = null_ref();
        t.A.ptr = t.A.rvec = t.A.cvec = t.B.ptr = t.B.rvec = t.B.cvec = null_ref();
        t.adam.w = t.adam.m = t.adam.v = t.adam.wt = t.adam.gexp = null_ref();
        for (auto &p : t.p) p = null_ref();
        return t;
    }
    void add(Task t, int n_tiles) {
        Stage &s = stages.back();
        t.tile_begin = s.n_tiles; t.n_tiles = n_tiles;
        s.n_tiles += n_tiles; s.task_end++;
        if (t.type == T_GEMM) has_gemm.back() = 1;
        tasks.push_back(t);
    }
    static Operand op(Ref p, int ld, int mn_major) {
        Operand o; memset(&o, 0, sizeof(o));
        o.ptr = p; o.ld = ld; o.mn_major = mn_major; o.xform = 0; o.rvec = o.cvec = null_ref();
        return o;
    }
    static Operand op_rank1(Ref h, int ld, int mn_major, Ref rvec, Ref cvec) {
        Operand o = op(h, ld, mn_major); o.xform = 1; o.rvec = rvec; o.cvec = cvec; return o;
    }
    void gemm(Operand a, Operand b, int M, int N, int K, Task t) {
        t.type = T_GEMM; t.A = a; t.B = b; t.M = M; t.N = N; t.K = K;
        t.tiles_m = cdiv(M, tm); t.tiles_n = cdiv(N, tn);
        add(t, t.tiles_m * t.tiles_n);
    }
    Task epi_bias_relu(Ref C, int ldc, Ref bias) { Task t = blank(T_GEMM); t.epi = EPI_BIAS_RELU; t.C = C; t.ldc = ldc; t.bias = bias; return t; }
    Task epi_bias(Ref C, int ldc, Ref bias) { Task t = blank(T_GEMM); t.epi = EPI_BIAS; t.C = C; t.ldc = ldc; t.bias = bias; return t; }
    Task epi_mask(Ref C, int ldc, Ref mask, int ldm) { Task t = blank(T_GEMM); t.epi = EPI_MASK; t.C = C; t.ldc = ldc; t.mask = mask; t.ld_mask = ldm; return t; }
    Task epi_store(Ref C, int ldc) { Task t = blank(T_GEMM); t.epi = EPI_STORE; t.C = C; t.ldc = ldc; return t; }

    // which optimizer a trainable net (0 policy, 1 q1, 2 q2) uses
    static int step_slot(int net) { return net == 0 ? SC_STEP_POLICY : (net == 1 ? SC_STEP_Q1 : SC_STEP_Q2); }
    bool apply() const { return key.dp_phase < 0; }
    bool exporting() const { return key.export_grads || key.dp_phase >= 0; }

    AdamArgs adam_args(int net, int64_t off_in_net) {
        AdamArgs a;
        a.w = A(L.param[net] + off_in_net); a.m = A(L.adam_m[net] + off_in_net); a.v = A(L.adam_v[net] + off_in_net);
        a.wt = net == 0 ? null_ref() : A(L.param[net + 2] + off_in_net);     // q1 -> q1_target, q2 -> q2_target
        a.gexp = exporting() ? A(L.grad[net] + off_in_net) : null_ref();
        a.step_slot = step_slot(net); a.apply = apply() ? 1 : 0;
        a.lr = h->cfg.lr; a.tau = h->cfg.tau;
        return a;
    }
    Task epi_adam(int net, int64_t off_in_net) { Task t = blank(T_GEMM); t.epi = EPI_ADAM; t.adam = adam_args(net, off_in_net); return t; }

    void bias_adam(int net, int64_t b_off, Operand dh, int Bn, int N) {
        Task t = blank(T_BIAS_ADAM);
        t.A = dh;
        AdamArgs a = adam_args(net, b_off);
        t.p[0] = a.w; t.p[1] = a.m; t.p[2] = a.v; t.p[3] = a.wt; t.p[4] = a.gexp;
        t.i[0] = Bn; t.i[1] = N; t.i[2] = a.step_slot; t.i[3] = a.apply; t.f[0] = a.lr; t.f[1] = a.tau;
        add(t, cdiv(N, 32));
    }

    void build() {
        const int B = key.B, H = L.hidden, nh = L.n_hidden, obs = L.obs, act = L.act, ldx = L.ldx, A2 = 2 * act, ldg = L.ldg;
        const NetLayout &P = L.pol, &Q = L.q;
        const Ref X2 = W(L.X), X1 = W(L.X + (int64_t)B * ldx), X3 = W(L.X + (int64_t)2 * B * ldx);
        const bool critics = key.dp_phase != 1, actor = key.dp_phase != 0;

        // ---- stage: gather ------------------------------------------------------------------------------------
        if (key.with_gather && critics) {
            begin_stage();
            Task t = blank(T_GATHER);
            t.p[0] = W(L.X); t.p[1] = W(L.r); t.p[2] = W(L.d);
            t.i[0] = B; t.i[1] = obs; t.i[2] = act; t.i[3] = ldx;
            add(t, cdiv(B, kThreads / 32));
        }
        if (critics) {
            // ---- policy forward on [s2 ; s] (M = 2B) + critic forward on (s,a), layer by layer ---------------------
            for (int l = 0; l < nh; l++) {
                begin_stage();
                const int in_p = P.in_of(l), in_q = Q.in_of(l);
                gemm(l == 0 ? op(X2, ldx, 0) : op(W(L.hp[l - 1]), H, 0), op(A(L.param[0] + P.w[l]), in_p, 0), 2 * B, H, in_p,
                     epi_bias_relu(W(L.hp[l]), H, A(L.param[0] + P.b[l])));
                for (int k = 0; k < 2; k++)
                    gemm(l == 0 ? op(X1, ldx, 0) : op(W(L.hc[k][l - 1]), H, 0), op(A(L.param[1 + k] + Q.w[l]), in_q, 0), B, H, in_q,
                         epi_bias_relu(W(L.hc[k][l]), H, A(L.param[1 + k] + Q.b[l])));
            }
            // ---- policy heads -> head_raw [2B, 2A] -----------------------------------------------------------------
            begin_stage();
            gemm(op(W(L.hp[nh - 1]), H, 0), op(A(L.param[0] + P.w_out), H, 0), 2 * B, A2, H,
                 epi_bias(W(L.head_raw), A2, A(L.param[0] + P.b_out)));
            // ---- reparameterised sample + log-prob for both batches -------------------------------------------------
            begin_stage();
            {
                Task t = blank(T_SAMPLE);
                t.p[0] = W(L.head_raw); t.p[1] = W(L.eps); t.p[2] = W(L.X); t.p[3] = W(L.logp);
                t.i[0] = B; t.i[1] = act; t.i[2] = obs; t.i[3] = ldx; t.i[4] = key.device_eps;
                t.f[0] = h->cfg.action_scale; t.f[1] = h->cfg.action_bias;
                add(t, cdiv(2 * B, kThreads / 32));
            }
            // ---- target critics on (s2, a2) ---------------------------------------------------------------------------
            for (int l = 0; l < nh; l++) {
                begin_stage();
                const int in_q = Q.in_of(l);
                for (int k = 0; k < 2; k++)
                    gemm(l == 0 ? op(X2, ldx, 0) : op(W(L.ht[k][l - 1]), H, 0), op(A(L.param[3 + k] + Q.w[l]), in_q, 0), B, H, in_q,
                         epi_bias_relu(W(L.ht[k][l]), H, A(L.param[3 + k] + Q.b[l])));
            }
            // ---- Bellman target, critic losses, dL/dq ---------------------------------------------------------------
            begin_stage();
            {
                Task t = blank(T_TARGET_LOSS);
                t.p[0] = W(L.ht[0][nh - 1]); t.p[1] = W(L.ht[1][nh - 1]); t.p[2] = W(L.hc[0][nh - 1]); t.p[3] = W(L.hc[1][nh - 1]);
                const int nets[4] = {3, 4, 1, 2};
                for (int k = 0; k < 4; k++) { t.p[4 + k] = A(L.param[nets[k]] + Q.w_out); t.p[8 + k] = A(L.param[nets[k]] + Q.b_out); }
                t.p[12] = W(L.r); t.p[13] = W(L.d); t.p[14] = W(L.logp); t.p[15] = key.use_isw ? W(L.isw) : null_ref();
                t.p[16] = W(L.y); t.p[17] = W(L.dq[0]); t.p[18] = W(L.dq[1]); t.p[19] = W(L.td);
                t.p[20] = W(L.wsnap[0]); t.p[21] = W(L.wsnap[1]); t.p[22] = W(L.loss_part);
                t.i[0] = B; t.i[1] = H; t.f[0] = h->cfg.gamma;
                add(t, cdiv(B, kThreads / 32));
            }
            // ---- critic backward.  dh of the last hidden layer is the implicit rank-1 operand dq * w_out * relu' ------
            //   stage s (1..nh): dX of layer l = nh-s (0-based, only while l >= 1) ; dW/db of layer l+1 ; last stage: dW/db of layer 0
            for (int s = 1; s <= nh; s++) {
                begin_stage();
                const int l = nh - s;          // layer whose dX is produced now (needs dh_l, writes dh_{l-1})
                for (int k = 0; k < 2; k++) {
                    const int net = 1 + k;
                    auto dh = [&](int layer, int mn_major) {   // dL/dh_layer as [B,H] operand
                        return layer == nh - 1 ? op_rank1(W(L.hc[k][nh - 1]), H, mn_major, W(L.dq[k]), W(L.wsnap[k]))
                                               : op(W(L.dhc[k][layer]), H, mn_major);
                    };
                    if (l >= 1)   // dh_{l-1} = (dh_l . W_l) * relu'(h_{l-1})
                        gemm(dh(l, 0), op(A(L.param[net] + Q.w[l]), Q.in_of(l), 1), B, Q.in_of(l), H,
                             epi_mask(W(L.dhc[k][l - 1]), H, W(L.hc[k][l - 1]), H));
                    auto dW = [&](int layer) {   // dW_layer = dh_layer^T . x_layer ; fused Adam + Polyak
                        const int in = Q.in_of(layer);
                        gemm(dh(layer, 1), layer == 0 ? op(X1, ldx, 1) : op(W(L.hc[k][layer - 1]), H, 1), H, in, B, epi_adam(net, Q.w[layer]));
                        bias_adam(net, Q.b[layer], dh(layer, 0), B, H);
                    };
                    if (l + 1 <= nh - 1) dW(l + 1);
                    if (s == nh) dW(0);
                    if (s == 1) {   // output layer (reads the live w_out; the rank-1 transforms read the snapshot)
                        Task t = blank(T_OUT_ADAM);
                        AdamArgs aw = adam_args(net, Q.w_out), ab = adam_args(net, Q.b_out);
                        t.p[0] = W(L.hc[k][nh - 1]); t.p[1] = W(L.dq[k]);
                        t.p[2] = aw.w; t.p[3] = aw.m; t.p[4] = aw.v; t.p[5] = aw.wt; t.p[6] = aw.gexp;
                        t.p[7] = ab.w; t.p[8] = ab.m; t.p[9] = ab.v; t.p[10] = ab.wt; t.p[11] = ab.gexp;
                        t.i[0] = B; t.i[1] = H; t.i[2] = aw.step_slot; t.i[3] = aw.apply; t.f[0] = aw.lr; t.f[1] = aw.tau;
                        add(t, cdiv(H, 32));
                    }
                }
            }
        }
        if (actor) {
            // ---- actor: updated critics on (s, a_new) ----------------------------------------------------------------
            for (int l = 0; l < nh; l++) {
                begin_stage();
                const int in_q = Q.in_of(l);
                for (int k = 0; k < 2; k++)
                    gemm(l == 0 ? op(X3, ldx, 0) : op(W(L.ha[k][l - 1]), H, 0), op(A(L.param[1 + k] + Q.w[l]), in_q, 0), B, H, in_q,
                         epi_bias_relu(W(L.ha[k][l]), H, A(L.param[1 + k] + Q.b[l])));
            }
            begin_stage();
            {
                Task t = blank(T_ACTOR_LOSS);
                t.p[0] = W(L.ha[0][nh - 1]); t.p[1] = W(L.ha[1][nh - 1]);
                t.p[2] = A(L.param[1] + Q.w_out); t.p[3] = A(L.param[2] + Q.w_out);
                t.p[4] = A(L.param[1] + Q.b_out); t.p[5] = A(L.param[2] + Q.b_out);
                t.p[6] = W(L.logp + B); t.p[7] = W(L.dqa[0]); t.p[8] = W(L.dqa[1]); t.p[9] = W(L.aloss_part);
                t.i[0] = B; t.i[1] = H; t.f[0] = -(float)act;
                add(t, cdiv(B, kThreads / 32));
            }
            // ---- dL/da through both critics (input gradients only: the Q weights are constants here, quirk Q2) -------
            for (int s = 1; s <= nh; s++) {
                begin_stage();
                const int l = nh - s;
                for (int k = 0; k < 2; k++) {
                    const int net = 1 + k;
                    Operand dh = l == nh - 1 ? op_rank1(W(L.ha[k][nh - 1]), H, 0, W(L.dqa[k]), A(L.param[net] + Q.w_out))
                                             : op(W(L.dha[k][l]), H, 0);
                    if (l >= 1)
                        gemm(dh, op(A(L.param[net] + Q.w[l]), H, 1), B, H, H, epi_mask(W(L.dha[k][l - 1]), H, W(L.ha[k][l - 1]), H));
                    else   // layer 0: only the action columns of W_0 [H, obs+act]
                        gemm(dh, op(A(L.param[net] + Q.w[0] + obs), obs + act, 1), B, act, H, epi_store(W(L.da[k]), act));
                }
            }
            begin_stage();
            {
                Task t = blank(T_SAMPLE_BWD);
                t.p[0] = W(L.da[0]); t.p[1] = W(L.da[1]); t.p[2] = W(L.head_raw + (int64_t)B * A2); t.p[3] = W(L.eps + (int64_t)B * act);
                t.p[4] = W(L.g_head);
                t.i[0] = B; t.i[1] = act; t.i[2] = ldg; t.f[0] = h->cfg.action_scale; t.f[1] = h->cfg.action_bias;
                add(t, cdiv(B * act, kThreads));
            }
            // ---- policy backward (current-state rows B..2B of the policy activations) --------------------------------
            //   stage 0: dh_{nh-1} from the heads ; stage s: dh_{nh-1-s}, dW of the layer above ; last: dW_0
            auto hp_cur = [&](int l) { return W(L.hp[l] + (int64_t)B * H); };
            for (int s = 0; s <= nh; s++) {
                begin_stage();
                if (s == 0) {
                    gemm(op(W(L.g_head), ldg, 0), op(A(L.param[0] + P.w_out), H, 1), B, H, A2,
                         epi_mask(W(L.dhp[nh - 1]), H, hp_cur(nh - 1), H));
                } else {
                    const int l = nh - s;      // dh_l available; produce dh_{l-1} (if l >= 1)
                    if (l >= 1)
                        gemm(op(W(L.dhp[l]), H, 0), op(A(L.param[0] + P.w[l]), H, 1), B, H, H,
                             epi_mask(W(L.dhp[l - 1]), H, hp_cur(l - 1), H));
                    if (s == 1) {              // heads: dW = g^T h_{nh-1}
                        gemm(op(W(L.g_head), ldg, 1), op(hp_cur(nh - 1), H, 1), A2, H, B, epi_adam(0, P.w_out));
                        bias_adam(0, P.b_out, op(W(L.g_head), ldg, 0), B, A2);
                    } else {
                        const int lw = l + 1;  // dW of the layer whose dX ran in the previous stage
                        gemm(op(W(L.dhp[lw]), H, 1), op(hp_cur(lw - 1), H, 1), H, H, B, epi_adam(0, P.w[lw]));
                        bias_adam(0, P.b[lw], op(W(L.dhp[lw]), H, 0), B, H);
                    }
                    if (s == nh) {
                        gemm(op(W(L.dhp[0]), H, 1), op(X1, ldx, 1), H, obs, B, epi_adam(0, P.w[0]));
                        bias_adam(0, P.b[0], op(W(L.dhp[0]), H, 0), B, H);
                    }
                }
            }
        }
        {   // own stage: the Adam epilogues of the previous stage still read the step counters
            begin_stage();
            Task t = blank(T_FINISH);
            const int ap = apply() ? 1 : 0;
            t.p[0] = critics ? W(L.loss_part) : null_ref();
            t.p[1] = actor ? W(L.aloss_part) : null_ref();
            t.p[2] = exporting() ? A(L.grad_scalars) : null_ref();
            t.i[0] = ap; t.i[1] = ap; t.i[2] = ap; t.i[3] = ap && h->cfg.auto_entropy; t.i[4] = ap;
            t.i[5] = cdiv(B, kThreads / 32); t.i[6] = h->cfg.auto_entropy; t.i[7] = ap;
            t.f[0] = h->cfg.lr; t.f[1] = (float)B;
            add(t, 1);
        }
    }
};

}  // namespace

void free_programs(sacb_handle h) {
    for (auto &kv : h->programs) {
        if (kv.second.graph) cudaGraphExecDestroy(kv.second.graph);
        cudaFree(kv.second.d_tasks); cudaFree(kv.second.d_stages);
    }
    h->programs.clear();
}

int check_error_flag(sacb_handle h) {
    int32_t flag = 0;
    SACB_CUDA(cudaMemcpyAsync(&flag, h->error_flag, sizeof(flag), cudaMemcpyDeviceToHost, h->stream));
    SACB_CUDA(cudaStreamSynchronize(h->stream));
    if (flag) {
        cudaMemsetAsync(h->error_flag, 0, sizeof(int32_t), h->stream);
        return fail(SACB_ERR_DEVICE, flag == 1 ? "tcgen05 pipeline watchdog: an mbarrier never completed"
                                              : "grid barrier watchdog: persistent kernel blocks were not co-resident");
    }
    return SACB_OK;
}

static int launch_range(sacb_handle h, ProgramInst &p, int s0, int s1, bool cooperative) {
    const bool tf32 = math_is_tc(h->cfg.math_mode);
    int needs_tc = 0, max_tiles = 1;
    for (int s = s0; s < s1; s++) { needs_tc |= p.stage_has_gemm[s]; max_tiles = std::max(max_tiles, p.stages[s].n_tiles * h->cfg.n_agents); }
    const size_t smem = math_smem(h->cfg.math_mode);
    uint64_t seed = h->cfg.seed;
    int tc_setup = tf32 ? needs_tc : 0;
    void *args[] = {(void *)&p.prog, (void *)&s0, (void *)&s1, (void *)&tc_setup, (void *)&seed};
    const void *fn = update_kernel_for(h->cfg.math_mode);
    if (cooperative) {
        const int grid = std::min(max_tiles, h->sm_count * h->coop_blocks_per_sm);
        SACB_CUDA(cudaLaunchCooperativeKernel(fn, dim3(grid), dim3(kThreads), args, smem, h->stream));
    } else {
        SACB_CUDA(cudaLaunchKernel(fn, dim3(max_tiles), dim3(kThreads), args, smem, h->stream));
    }
    return SACB_OK;
}

static int record_step(sacb_handle h, ProgramInst &p) {
    if (h->cfg.launch_mode == SACB_LAUNCH_PERSISTENT) {
        SACB_CUDA(cudaMemsetAsync(h->barrier, 0, sizeof(unsigned int), h->stream));
        return launch_range(h, p, 0, (int)p.stages.size(), true);
    }
    for (int s = 0; s < (int)p.stages.size(); s++) {
        int rc = launch_range(h, p, s, s + 1, false);
        if (rc) return rc;
    }
    return SACB_OK;
}

int get_program(sacb_handle h, const ProgramKey &key, ProgramInst **out) {
    auto it = h->programs.find(key);
    if (it != h->programs.end()) { *out = &it->second; return SACB_OK; }
    if (key.B < 1 || key.B > h->cfg.max_batch) return fail(SACB_ERR_ARG, "batch size exceeds max_batch of the handle");
    Builder b(h, key);
    b.build();
    ProgramInst &p = h->programs[key];
    p.tasks = b.tasks; p.stages = b.stages; p.stage_has_gemm = b.has_gemm;
    for (auto &s : p.stages) { p.n_tiles_total += s.n_tiles; p.max_stage_tiles = std::max(p.max_stage_tiles, s.n_tiles); }
    SACB_CUDA(cudaMalloc(&p.d_tasks, p.tasks.size() * sizeof(Task)));
    SACB_CUDA(cudaMalloc(&p.d_stages, p.stages.size() * sizeof(Stage)));
    SACB_CUDA(cudaMemcpyAsync(p.d_tasks, p.tasks.data(), p.tasks.size() * sizeof(Task), cudaMemcpyHostToDevice, h->stream));
    SACB_CUDA(cudaMemcpyAsync(p.d_stages, p.stages.data(), p.stages.size() * sizeof(Stage), cudaMemcpyHostToDevice, h->stream));
    SACB_CUDA(cudaStreamSynchronize(h->stream));
    Program &P = p.prog;
    P.tasks = p.d_tasks; P.stages = p.d_stages; P.n_stages = (int)p.stages.size(); P.n_agents = h->cfg.n_agents;
    P.bases.arena = h->arena; P.bases.ws = h->ws; P.bases.arena_stride = h->L.arena_size; P.bases.ws_stride = h->L.ws_size;
    P.scalars = make_ref(0, h->L.scalars);
    P.barrier = h->barrier;
    P.ring = h->ring; P.ring_agent_stride = h->cfg.capacity * h->ring_row; P.ring_row = (int32_t)h->ring_row;
    P.slots = h->slots; P.slots_stride = h->cfg.max_batch;
    P.error_flag = h->error_flag;
    P.trace = nullptr;
    p.kernels_per_step = h->cfg.launch_mode == SACB_LAUNCH_PERSISTENT ? 1 : (int)p.stages.size();

    // capture one step into a CUDA graph (stage kernels, or memset + the single cooperative launch)
    cudaGraph_t graph = nullptr;
    cudaError_t e = cudaStreamBeginCapture(h->stream, cudaStreamCaptureModeThreadLocal);
    if (e == cudaSuccess) {
        int rc = record_step(h, p);
        cudaError_t e2 = cudaStreamEndCapture(h->stream, &graph);
        if (rc == SACB_OK && e2 == cudaSuccess && graph) {
            if (cudaGraphInstantiate(&p.graph, graph, 0) != cudaSuccess) p.graph = nullptr;
        }
        if (graph) cudaGraphDestroy(graph);
    }
    cudaGetLastError();   // a failed capture falls back to direct launches
    *out = &p;
    return SACB_OK;
}

int launch_program(sacb_handle h, ProgramInst &p) {
    h->kernel_launches += p.kernels_per_step;
    if (p.graph) { SACB_CUDA(cudaGraphLaunch(p.graph, h->stream)); return SACB_OK; }
    return record_step(h, p);
}

}  // namespace sacb

// ================================================================================================================
// instrumentation entry points that need the kernel symbol
// ================================================================================================================
using namespace sacb;

extern "C" int sacb_time_stages(sacb_handle h, int64_t B, float *us_out, int cap) {
    if (!h) return fail(SACB_ERR_ARG, "null handle");
    ProgramKey key{(int)B, 0, 0, 1, 0, -1};
    ProgramInst *p;
    int rc = get_program(h, key, &p);
    if (rc) return rc;
    const bool trace = getenv("SACB_TRACE") != nullptr;
    unsigned long long *d_trace = nullptr;
    std::vector<unsigned long long> h_trace;
    if (trace) { SACB_CUDA(cudaMalloc(&d_trace, sizeof(unsigned long long) * 8 * 4096)); p->prog.trace = d_trace; h_trace.resize(8 * 4096); }
    const int n = std::min<int>(cap, (int)p->stages.size());
    std::vector<cudaEvent_t> ev(p->stages.size() + 1);
    for (auto &e : ev) cudaEventCreate(&e);
    std::vector<float> acc(p->stages.size(), 0.f);
    const int reps = 20;
    for (int r = 0; r < reps + 3; r++) {
        cudaEventRecord(ev[0], h->stream);
        for (int s = 0; s < (int)p->stages.size(); s++) {
            // force the staged form regardless of launch_mode: this is a per-stage profile
            const bool tf32 = math_is_tc(h->cfg.math_mode);
            int s0 = s, s1 = s + 1, tc_setup = tf32 ? p->stage_has_gemm[s] : 0;
            uint64_t seed = h->cfg.seed;
            void *args[] = {(void *)&p->prog, (void *)&s0, (void *)&s1, (void *)&tc_setup, (void *)&seed};
            SACB_CUDA(cudaLaunchKernel(update_kernel_for(h->cfg.math_mode), dim3(std::max(1, p->stages[s].n_tiles * h->cfg.n_agents)), dim3(kThreads), args,
                                       math_smem(h->cfg.math_mode), h->stream));
            cudaEventRecord(ev[s + 1], h->stream);
            if (trace && r == reps + 2) {
                SACB_CUDA(cudaStreamSynchronize(h->stream));
                const int grid = std::max(1, p->stages[s].n_tiles * h->cfg.n_agents);
                SACB_CUDA(cudaMemcpy(h_trace.data(), d_trace, sizeof(unsigned long long) * 8 * std::min(grid, 4096), cudaMemcpyDeviceToHost));
                double a[8] = {0}; unsigned long long t0 = ~0ull, t5 = 0;
                const int g = std::min(grid, 4096);
                for (int b = 0; b < g; b++) { t0 = std::min(t0, h_trace[b * 8]); t5 = std::max(t5, h_trace[b * 8 + 5]);
                    a[1] += (double)(h_trace[b * 8 + 1] - h_trace[b * 8]); a[2] += (double)(h_trace[b * 8 + 2] - h_trace[b * 8 + 1]);
                    a[3] += (double)(h_trace[b * 8 + 3] - h_trace[b * 8 + 2]); a[4] += (double)(h_trace[b * 8 + 4] - h_trace[b * 8 + 1]);
                    a[5] += (double)(h_trace[b * 8 + 5] - h_trace[b * 8 + 4]); }
                fprintf(stderr, "[trace] stage %2d grid %4d has_gemm %d  span %.2f us | per-CTA avg: setup %.2f  mainloop %.2f  epilogue %.2f  tile %.2f  teardown %.2f us\n",
                        s, grid, p->stage_has_gemm[s], (t5 - t0) * 1e-3, a[1] / g * 1e-3, a[2] / g * 1e-3, a[3] / g * 1e-3, a[4] / g * 1e-3, a[5] / g * 1e-3);
            }
        }
        SACB_CUDA(cudaStreamSynchronize(h->stream));
        if (r >= 3)
            for (int s = 0; s < (int)p->stages.size(); s++) { float ms; cudaEventElapsedTime(&ms, ev[s], ev[s + 1]); acc[s] += ms * 1000.f / reps; }
    }
    for (int s = 0; s < n; s++) us_out[s] = acc[s];
    if (trace) { p->prog.trace = nullptr; cudaFree(d_trace); }
    for (auto &e : ev) cudaEventDestroy(e);
    h->kernel_launches += (int64_t)(reps + 3) * p->stages.size();
    return (int)p->stages.size();
}

namespace sacb {
const void *update_kernel_for(int m) {
    return m == SACB_MATH_TF32X3 ? (const void *)sac_update_kernel<SACB_MATH_TF32X3>
         : m == SACB_MATH_TF32   ? (const void *)sac_update_kernel<SACB_MATH_TF32>
                                 : (const void *)sac_update_kernel<SACB_MATH_FP32>;
}
int init_kernel_attributes(sacb_handle h) {
    const int m = h->cfg.math_mode;
    if (m < 0 || m > SACB_MATH_TF32X3) return fail(SACB_ERR_ARG, "bad math_mode");
    if (math_is_tc(m) && (h->cfg.hidden_dim > tc::kXkMax || h->cfg.max_batch > tc::kXkMax)) return fail(SACB_ERR_ARG, "tensor-core modes need hidden_dim, max_batch <= 2048");
    SACB_CUDA(cudaFuncSetAttribute(update_kernel_for(m), cudaFuncAttributeMaxDynamicSharedMemorySize, (int)math_smem(m)));
    int nb = 0;
    SACB_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, update_kernel_for(m), kThreads, math_smem(m)));
    h->coop_blocks_per_sm = std::max(1, std::min(nb, math_is_tc(m) ? 2 : 4));
    return SACB_OK;
}
}  // namespace sacb
