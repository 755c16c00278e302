// Large-batch data-parallel mode (BASELINE.json configs[3]): each rank runs the backward half of a phase on its
// B_local rows with gradients exported (not applied), the host all-reduces the gradient arena over NCCL
// (torch.distributed), then sacb_dp_apply runs Adam (+ Polyak) on the averaged gradients.
//   phase 0 = critics (sac_imp.py:101-113), phase 1 = actor + temperature (sac_imp.py:116-135)
#include "handle.h"
#include "gemm.cuh"

namespace sacb {

__global__ void dp_apply_kernel(float *w, float *m, float *v, float *wt, const float *g, int64_t n, float scale, const float *scalars, int step_slot, float lr, float tau) {
    float ss, bs;
    adam_factors_cached(scalars, step_slot, ss, bs);
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
        adam_element(g[i] * scale, w + i, m + i, v + i, wt ? wt + i : nullptr, nullptr, 1, ss, bs, tau);
}

__global__ void dp_finish_kernel(float *scalars, const float *g_scalars, int phase, int auto_entropy, const float2 *adam_table) {
    if (threadIdx.x || blockIdx.x) return;
    if (phase == 0) {
        for (int slot = SC_STEP_Q1; slot <= SC_STEP_Q2; slot++) {
            const int step = __float_as_int(scalars[slot]) + 1;
            scalars[slot] = __int_as_float(step);
            adam_factors_store(scalars, slot, step, adam_table);
        }
    } else {
        const int n_upd = __float_as_int(scalars[SC_N_UPDATES]);
        float alpha_next = scalars[SC_ALPHA0 + (n_upd & 1)];
        if (auto_entropy) {
            float ss, bs;
            adam_factors_cached(scalars, SC_STEP_ALPHA, ss, bs);
            adam_element(g_scalars[0], &scalars[SC_LOG_ALPHA], &scalars[SC_LOG_ALPHA_M], &scalars[SC_LOG_ALPHA_V], nullptr, nullptr, 1, ss, bs, 0.f);
            alpha_next = expf(scalars[SC_LOG_ALPHA]);
            const int step = __float_as_int(scalars[SC_STEP_ALPHA]) + 1;
            scalars[SC_STEP_ALPHA] = __int_as_float(step);
            adam_factors_store(scalars, SC_STEP_ALPHA, step, adam_table);
        }
        scalars[SC_ALPHA0 + ((n_upd + 1) & 1)] = alpha_next;
        const int pstep = __float_as_int(scalars[SC_STEP_POLICY]) + 1;
        scalars[SC_STEP_POLICY] = __int_as_float(pstep);
        adam_factors_store(scalars, SC_STEP_POLICY, pstep, adam_table);
        scalars[SC_N_UPDATES] = __int_as_float(n_upd + 1);
    }
}

}  // namespace sacb
using namespace sacb;

// gradients of a phase are contiguous in the arena: phase 0 = grad[q1] | grad[q2], phase 1 = grad[policy] (+ log_alpha grad kept
// in the scalar-gradient block right after grad[q2]; the host reduces it as a second tiny buffer)
extern "C" int sacb_dp_grad_buffer(sacb_handle h, int phase, void **dev_ptr, int64_t *n_floats) {
    if (!h || !dev_ptr || !n_floats || phase < 0 || phase > 2) return fail(SACB_ERR_ARG, "bad argument");
    if (phase == 0) { *dev_ptr = h->arena + h->L.grad[1]; *n_floats = 2 * h->L.q.size; }
    else if (phase == 1) { *dev_ptr = h->arena + h->L.grad[0]; *n_floats = h->L.pol.size; }
    else { *dev_ptr = h->arena + h->L.grad_scalars; *n_floats = 32; }
    return SACB_OK;
}

namespace sacb { int replay_stage_slots(sacb_handle h, const int64_t *idx, int64_t B); }

extern "C" int sacb_dp_backward(sacb_handle h, int phase, int64_t B_local, const int64_t *idx, const float *eps_next, const float *eps_cur) {
    if (!h || phase < 0 || phase > 1 || h->cfg.n_agents != 1) return fail(SACB_ERR_ARG, "bad argument");
    if (B_local < 1 || B_local > h->L.maxB) return fail(SACB_ERR_ARG, "batch size out of range");
    if ((eps_next == nullptr) != (eps_cur == nullptr)) return fail(SACB_ERR_ARG, "pass both eps arrays or neither");
    int rc;
    // every rank averages over its own B_local rows; with equal B_local the all-reduced mean of the rank means is the global mean
    if (phase == 0) {
        if (idx && (rc = replay_stage_slots(h, idx, B_local))) return rc;
        if (eps_next) {
            const int64_t n = B_local * h->cfg.act_dim;
            SACB_CUDA(cudaMemcpyAsync(h->ws + h->L.eps, eps_next, sizeof(float) * n, cudaMemcpyHostToDevice, h->stream));
            SACB_CUDA(cudaMemcpyAsync(h->ws + h->L.eps + n, eps_cur, sizeof(float) * n, cudaMemcpyHostToDevice, h->stream));
            h->dp_device_eps = 0;
        } else {
            h->dp_device_eps = 1;
        }
    }
    ProgramKey key{(int)B_local, phase == 0 ? 1 : 0, 1, h->dp_device_eps, 0, phase};
    ProgramInst *p;
    rc = get_program(h, key, &p);
    if (rc) return rc;
    return launch_program(h, *p);
}

extern "C" int sacb_dp_apply(sacb_handle h, int phase) {
    if (!h || phase < 0 || phase > 1) return fail(SACB_ERR_ARG, "bad argument");
    const Layout &L = h->L;
    float *ar = h->arena;
    float *sc = ar + L.scalars;
    auto run = [&](int net, int64_t size) {
        dp_apply_kernel<<<296, 256, 0, h->stream>>>(ar + L.param[net], ar + L.adam_m[net], ar + L.adam_v[net], net ? ar + L.param[net + 2] : nullptr, ar + L.grad[net], size,
                                                    1.0f, sc, net == 0 ? SC_STEP_POLICY : (net == 1 ? SC_STEP_Q1 : SC_STEP_Q2), h->cfg.lr, h->cfg.tau);
        h->kernel_launches++;
    };
    h->shadows_valid = false;      // the element-wise apply writes fp32 weights only: the next program re-derives the shadows
    if (phase == 0) { run(1, L.q.size); run(2, L.q.size); } else run(0, L.pol.size);
    dp_finish_kernel<<<1, 32, 0, h->stream>>>(sc, ar + L.grad_scalars, phase, h->cfg.auto_entropy, h->adam_table);
    h->kernel_launches++;
    SACB_CUDA(cudaGetLastError());
    return SACB_OK;
}
