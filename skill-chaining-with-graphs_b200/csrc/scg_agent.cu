// scg_agent.cu - the fused lock-step agent step: K1 (env step) -> K2+K4 (control) -> K3 (traces).
//
// Mirrors oracle/agent.py SkillChainAgent.step, numbered steps 1-8 of its docstring (the reference
// has no code: /root/reference/README.md:1-2).  One C call launches, on one stream:
//   k_step            s2, r_env, flags = env.step(a)                       (scg_step.cu)
//   k_agent_control   initiation bits of s2 (K4), termination, option reward, Q_o(s, a) and
//                     Q_o(s2, .) with shared weight loads (K2), eps-greedy a2, TD error, the
//                     48-byte update record for K3, example-ring append, env reset, option
//                     re-selection + first action under the new option
//   k_trace, k_reduce the trace sweep and dW reduction                     (scg_sarsa.cu)
// Weight application (and the cross-GPU allreduce of dW / cnt) happens outside, every sync interval.
#include <algorithm>

#include "scg_common.cuh"

int scg_launch_step(const scg_map_t *map, int B, const float *x, const float *y, const float *vx, const float *vy,
                    const int *action, float *x2, float *y2, float *vx2, float *vy2, float *reward, int *flags,
                    int cull, cudaStream_t st);
int scg_launch_trace(scg_ctx *ctx, int B, const float *rec, float *trace, float gl, float *dW, cudaStream_t st,
                     cudaEvent_t *ev2 = nullptr);

struct ControlArgs {
    scg_agent_t ag;
    const unsigned char *map_blob;
};

template <int N1>
__global__ void __launch_bounds__(32 * N1) k_agent_control(const __grid_constant__ ControlArgs args) {
    constexpr int F = N1 * N1 * N1 * N1;
    __shared__ float zsh[2][4][2][32];           // [state][dim][cos, sin][lane]
    __shared__ float part[N1][2 * SCG_A][32];    // [warp][q value][lane]
    const scg_agent_t &g = args.ag;
    const ScgMapHeader *mh = reinterpret_cast<const ScgMapHeader *>(args.map_blob);
    const int K = g.K;
    const int gest = min(g.n_active, K - 1);
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    for (int base = blockIdx.x * 32; base < g.B; base += gridDim.x * 32) {
        const bool valid = base + lane < g.B;
        const int b = valid ? base + lane : g.B - 1;
        float sx = g.x[b], sy = g.y[b], svx = g.vx[b], svy = g.vy[b];
        float nx = g.x2[b], ny = g.y2[b], nvx = g.vx2[b], nvy = g.vy2[b];
        const int o = g.option[b];
        // phasors of s and s2: warp w evaluates dimensions w, w + N1, ... and shares them
        {
            float sa[4], sb[4];
            scg_normalise(sx, sy, svx, svy, sa);
            scg_normalise(nx, ny, nvx, nvy, sb);
#pragma unroll
            for (int d = 0; d < 4; ++d) {
                if (d % N1 == w) {
                    float sn, cs;
                    sincospif(sa[d], &sn, &cs);
                    zsh[0][d][0][lane] = cs; zsh[0][d][1][lane] = sn;
                    sincospif(sb[d], &sn, &cs);
                    zsh[1][d][0][lane] = cs; zsh[1][d][1][lane] = sn;
                }
            }
        }
        __syncthreads();
        float2 za[4], zb[4];
#pragma unroll
        for (int d = 0; d < 4; ++d) {
            za[d] = make_float2(zsh[0][d][0][lane], zsh[0][d][1][lane]);
            zb[d] = make_float2(zsh[1][d][0][lane], zsh[1][d][1][lane]);
        }
        // 4a: this warp's share of Q_o(s, .) and Q_o(s2, .)
        {
            float qa[SCG_A], qb[SCG_A];
            scg_q_pair_c0<N1>(w, za, zb, g.Wt + (size_t)o * F * SCG_WT_STRIDE, qa, qb);
#pragma unroll
            for (int i = 0; i < SCG_A; ++i) {
                part[w][i][lane] = qa[i];
                part[w][SCG_A + i][lane] = qb[i];
            }
        }
        __syncthreads();
        if (w == 0 && valid) {
            float qa[SCG_A], qb[SCG_A];
#pragma unroll
            for (int i = 0; i < SCG_A; ++i) {
                float s0 = part[0][i][lane], s1 = part[0][SCG_A + i][lane];
#pragma unroll
                for (int ww = 1; ww < N1; ++ww) {
                    s0 += part[ww][i][lane];
                    s1 += part[ww][SCG_A + i][lane];
                }
                qa[i] = s0; qb[i] = s1;
            }
            const uint32_t env = g.env_offset + (uint32_t)b;
            const int a = g.action[b];
            const float r_env = g.reward[b];
            const bool env_done = (g.flags[b] & SCG_FLAG_DONE) != 0;
            // 2-3: initiation bits of s2, termination, option reward
            uint32_t bits = scg_init_bits(g.theta, K, g.active_mask, nx, ny);
            uint32_t pm = g.parents[o];
            bool hit = (((pm & SCG_GOAL_BIT) != 0) && env_done) || ((bits & pm & ~SCG_GOAL_BIT) != 0);
            int t_opt = g.t_opt[b] + 1, ep = g.ep_steps[b] + 1;
            bool left = ((g.active_mask >> o) & 1u) && !((bits >> o) & 1u);
            bool ep_timeout = (ep >= g.max_episode_steps) && !env_done;
            bool term = env_done || hit || (t_opt >= g.option_timeout) || left || ep_timeout;
            float r = __fadd_rn(r_env, (hit && !env_done) ? g.option_bonus : 0.f);
            // 4b: a2, TD error
            int a2 = scg_eps_greedy(qb, g.epsilon, scg_draw(g.seed, env, g.step, SCG_STREAM_ACTION));
            float qsa = 0.f, qs2 = 0.f;
#pragma unroll
            for (int i = 0; i < SCG_A; ++i) {
                qsa = (i == a) ? qa[i] : qsa;
                qs2 = (i == a2) ? qb[i] : qs2;
            }
            float delta = __fsub_rn(__fadd_rn(r, __fmul_rn(__fmul_rn(g.gamma, term ? 0.f : 1.f), qs2)), qsa);
            g.delta[b] = delta;
            // 5: hand the update to K3
            float4 *rec = reinterpret_cast<float4 *>(g.rec) + (size_t)b * 3;
            uint32_t meta = (uint32_t)a | ((uint32_t)o << 8) | (term ? SCG_META_ZERO_AFTER : 0u) | SCG_META_ACTIVE;
            rec[0] = make_float4(za[0].x, za[0].y, za[1].x, za[1].y);
            rec[1] = make_float4(za[2].x, za[2].y, za[3].x, za[3].y);
            rec[2] = make_float4(delta, __uint_as_float(meta), 0.f, 0.f);
            {   // cnt[o] += 1, one atomic per (warp, option)
                uint32_t peers = __match_any_sync(__activemask(), o);
                if ((__ffs(peers) - 1) == lane) atomicAdd(g.cnt + o, __popc(peers));
            }
            float ret = g.ep_return[b] + r_env;
            // 6: example for option o's initiation classifier
            if (term) {
                int eslot = atomicAdd(g.ex_count + o, 1) % (int)g.example_capacity;
                size_t ei = (size_t)o * g.example_capacity + eslot;
                g.ex_xy[2 * ei] = g.start_xy[2 * b];
                g.ex_xy[2 * ei + 1] = g.start_xy[2 * b + 1];
                g.ex_label[ei] = hit ? 1 : 0;
                atomicAdd((hit ? g.n_success : g.n_fail) + o, 1);
            }
            // 7: env reset
            bool reset = env_done || ep_timeout;
            if (reset) {
                uint4 rr = scg_draw(g.seed, env, g.step, SCG_STREAM_RESET);
                int ns = mh->n_starts;
                int pick = min((int)__fmul_rn(scg_u01(rr.x), (float)ns), ns - 1);
                const float2 *starts = reinterpret_cast<const float2 *>(args.map_blob + mh->off_starts);
                float2 s0 = starts[pick];
                nx = s0.x; ny = s0.y; nvx = 0.f; nvy = 0.f;
                atomicAdd(g.stats + 0, 1);
                if (env_done) atomicAdd(g.stats + 1, 1);
                atomicAdd(reinterpret_cast<float *>(g.stats + 2), ret);
                ret = 0.f;
                ep = 0;
                g.x2[b] = nx; g.y2[b] = ny; g.vx2[b] = nvx; g.vy2[b] = nvy;
            }
            g.ep_return[b] = ret;
            g.ep_steps[b] = ep;
            // 8: option re-selection (rare: this lane walks all F features of the new option alone)
            int o_next = o, a_next = a2;
            if (term) {
                uint32_t bn = reset ? scg_init_bits(g.theta, K, g.active_mask, nx, ny) : bits;
                o_next = bn ? (__ffs(bn) - 1) : gest;
                float2 zn[4];
                if (reset) scg_phasors(nx, ny, nvx, nvy, zn);
                else { zn[0] = zb[0]; zn[1] = zb[1]; zn[2] = zb[2]; zn[3] = zb[3]; }
                float qn[SCG_A];
                scg_q_one<N1>(zn, g.Wt + (size_t)o_next * F * SCG_WT_STRIDE, qn);
                a_next = scg_eps_greedy(qn, g.epsilon, scg_draw(g.seed, env, g.step, SCG_STREAM_RESELECT));
                t_opt = 0;
                g.start_xy[2 * b] = nx;
                g.start_xy[2 * b + 1] = ny;
            }
            g.t_opt[b] = t_opt;
            g.option[b] = o_next;
            g.action[b] = a_next;
        }
        __syncthreads();
    }
}

extern "C" int scg_agent_step(const scg_map_t *map, scg_ctx_t *ctx, const scg_agent_t *ag, void *stream) {
    if (!map || !ctx || !ag) return SCG_EINVAL;
    if (ag->K != ctx->K || ag->order != ctx->order) return SCG_EINVAL;
    if (ag->K < 1 || ag->K > SCG_MAX_OPTIONS) return SCG_ELIMIT;
    if (ag->B <= 0) return ag->B == 0 ? 0 : SCG_EINVAL;
    cudaStream_t st = (cudaStream_t)stream;
    cudaEvent_t *ev = nullptr;
    if (ctx->prof_on && ctx->prof_n < ctx->prof_cap) ev = ctx->prof_ev + (size_t)5 * ctx->prof_n++;
    if (ev) SCG_CUDA_OK(cudaEventRecord(ev[0], st));
    int rc = scg_launch_step(map, ag->B, ag->x, ag->y, ag->vx, ag->vy, ag->action, ag->x2, ag->y2, ag->vx2, ag->vy2,
                             ag->reward, ag->flags, ag->cull, st);
    if (rc) return rc;
    if (ev) SCG_CUDA_OK(cudaEventRecord(ev[1], st));
    ControlArgs args;
    args.ag = *ag;
    args.map_blob = map->d_blob;
    int grid = std::max(1, std::min((ag->B + 31) / 32, SCG_NUM_SMS * 32));
    DISPATCH_ORDER(ag->order, k_agent_control<N1><<<grid, 32 * N1, 0, st>>>(args));
    SCG_LAUNCH_CHECK();
    if (ev) SCG_CUDA_OK(cudaEventRecord(ev[2], st));
    return scg_launch_trace(ctx, ag->B, ag->rec, ag->trace, ag->gamma * ag->lambda, ag->dW, st, ev ? ev + 3 : nullptr);
}

extern "C" int scg_agent_step_host(const scg_map_t *map, scg_ctx_t *ctx, const scg_agent_t *ag,
                                   const float *h_state_soa, const int *h_action, float *h_state2_soa,
                                   float *h_reward, int *h_flags, int *h_action2, float *h_delta, void *stream) {
    if (!map || !ctx || !ag || !h_state_soa || !h_action || !h_state2_soa || !h_reward || !h_flags || !h_action2 ||
        !h_delta)
        return SCG_EINVAL;
    cudaStream_t st = (cudaStream_t)stream;
    size_t n = (size_t)ag->B * sizeof(float);
    float *in[4] = {ag->x, ag->y, ag->vx, ag->vy}, *out[4] = {ag->x2, ag->y2, ag->vx2, ag->vy2};
    for (int i = 0; i < 4; ++i)
        SCG_CUDA_OK(cudaMemcpyAsync(in[i], h_state_soa + (size_t)i * ag->B, n, cudaMemcpyHostToDevice, st));
    SCG_CUDA_OK(cudaMemcpyAsync(ag->action, h_action, n, cudaMemcpyHostToDevice, st));
    int rc = scg_agent_step(map, ctx, ag, stream);
    if (rc) return rc;
    for (int i = 0; i < 4; ++i)
        SCG_CUDA_OK(cudaMemcpyAsync(h_state2_soa + (size_t)i * ag->B, out[i], n, cudaMemcpyDeviceToHost, st));
    SCG_CUDA_OK(cudaMemcpyAsync(h_reward, ag->reward, n, cudaMemcpyDeviceToHost, st));
    SCG_CUDA_OK(cudaMemcpyAsync(h_flags, ag->flags, n, cudaMemcpyDeviceToHost, st));
    SCG_CUDA_OK(cudaMemcpyAsync(h_action2, ag->action, n, cudaMemcpyDeviceToHost, st));
    SCG_CUDA_OK(cudaMemcpyAsync(h_delta, ag->delta, n, cudaMemcpyDeviceToHost, st));
    SCG_CUDA_OK(cudaStreamSynchronize(st));
    return 0;
}

extern "C" int scg_profile_begin(scg_ctx_t *ctx, int max_steps) {
    if (!ctx || max_steps <= 0) return SCG_EINVAL;
    if (max_steps > ctx->prof_cap) {
        for (int i = 0; i < 5 * ctx->prof_cap; ++i) cudaEventDestroy(ctx->prof_ev[i]);
        free(ctx->prof_ev);
        ctx->prof_cap = 0;
        ctx->prof_ev = (cudaEvent_t *)calloc((size_t)5 * max_steps, sizeof(cudaEvent_t));
        if (!ctx->prof_ev) return SCG_ENOMEM;
        for (int i = 0; i < 5 * max_steps; ++i) SCG_CUDA_OK(cudaEventCreate(&ctx->prof_ev[i]));
        ctx->prof_cap = max_steps;
    }
    ctx->prof_n = 0;
    ctx->prof_on = 1;
    return 0;
}

extern "C" int scg_profile_end(scg_ctx_t *ctx, float *ms, int *steps) {
    if (!ctx || !ms || !steps) return SCG_EINVAL;
    ctx->prof_on = 0;
    for (int k = 0; k < 4; ++k) ms[k] = 0.f;
    *steps = ctx->prof_n;
    for (int i = 0; i < ctx->prof_n; ++i) {
        cudaEvent_t *ev = ctx->prof_ev + (size_t)5 * i;
        SCG_CUDA_OK(cudaEventSynchronize(ev[4]));
        for (int k = 0; k < 4; ++k) {
            float t = 0.f;
            SCG_CUDA_OK(cudaEventElapsedTime(&t, ev[k], ev[k + 1]));
            ms[k] += t;
        }
    }
    return 0;
}

extern "C" void scg_agent_swap(scg_agent_t *ag) {
    if (!ag) return;
    std::swap(ag->x, ag->x2); std::swap(ag->y, ag->y2);
    std::swap(ag->vx, ag->vx2); std::swap(ag->vy, ag->vy2);
}
