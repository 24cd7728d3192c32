// scg_ctl.cu - the option-creation controller on the device (sm_100a): deterministic example rings and the
// promote-and-fit kernel.
//
// Mirrors oracle/agent.py SkillChainAgent.step item 6 (example ring append) and SkillChainAgent.manage (the reference
// has no code: /root/reference/README.md:1-2).
//
// k_ring   The step kernel leaves one event byte per env-step (option terminated? hit? which option) and the option's
//          start position in the step record.  k_ring turns the events of the steps not yet processed into ring
//          appends in exactly the oracle's order - steps in order, envs in order within a step - so the ring contents
//          (including which examples survive once a ring wraps) do not depend on launch geometry or atomic timing:
//          per-CTA per-option counts -> grid barrier -> exclusive prefix -> slot = (ex_count + rank) % capacity, and
//          only the last `capacity` new examples of an option are written (no two writers per slot).
// k_manage One CTA.  Reads the gestating option's success count, and if it qualifies fits its logistic initiation
//          classifier on its ring by batch gradient descent from theta = 0, flips the option to active and wires the
//          next gestating slot's parents - all in device memory, so the agent loop never waits for the host.  Across
//          ranks the per-step gradient sums and example counts travel through NVLink peer memory (one flag round per
//          gradient step) and are added in rank order: every rank ends with the bit-identical theta of one fit on the
//          union of all ranks' examples.
#include <stdlib.h>
#include <string.h>

#include <algorithm>

#include "scg_common.cuh"
#include "scg_xchg.cuh"

int scg_prof_push(scg_ctx *ctx, int kind, cudaStream_t st, bool end);

#define RING_CTAS 128
#define RING_NT 1024
#define RING_TILE (RING_NT * 4)      // bytes per tile: one 32-bit word per thread, coalesced

// scratch layout (unsigned int): [RING_CTAS][16] per-CTA option counts | barrier counter | done ticket
#define RING_BAR (RING_CTAS * SCG_MAX_OPTIONS)
#define RING_DONE (RING_BAR + 1)
#define RING_WORDS (RING_DONE + 1)

// exclusive prefix of one value per thread over the CTA (1024 threads); *total receives the sum
__device__ __forceinline__ uint32_t ring_block_scan(uint32_t v, uint32_t *warp_sums, uint32_t *total) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    uint32_t inc = v;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const uint32_t u = __shfl_up_sync(0xffffffffu, inc, d);
        if (lane >= d) inc += u;
    }
    if (lane == 31) warp_sums[warp] = inc;
    __syncthreads();
    if (warp == 0) {
        const uint32_t w = warp_sums[lane];
        uint32_t winc = w;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const uint32_t u = __shfl_up_sync(0xffffffffu, winc, d);
            if (lane >= d) winc += u;
        }
        warp_sums[lane] = winc - w;
        if (lane == 31) *total = winc;
    }
    __syncthreads();
    return warp_sums[warp] + inc - v;
}

// CTA c owns the contiguous tiles [c * tiles_per_cta, ...) of the flat [step][env] event-byte array (n bytes, n % 4 == 0
// after padding by the caller's layout: the array is read as aligned 32-bit words, `mis` leading bytes are skipped).
__global__ void __launch_bounds__(RING_NT) k_ring(int n, int K, const uint8_t *__restrict__ ev, const float2 *__restrict__ pos,
                                                  float *ex_xy, uint8_t *ex_label, long long *ex_count, uint32_t cap,
                                                  unsigned int *scratch, unsigned int target, int tiles_per_cta) {
    __shared__ uint32_t lst[RING_TILE];                 // this tile's events in flat order: option | hit << 4 | byte index << 8
    __shared__ uint32_t rk[RING_TILE];                  // their rank among the CTA's events of the same option
    __shared__ uint32_t warp_sums[32];
    __shared__ uint32_t cnt_s[SCG_MAX_OPTIONS], base_s[SCG_MAX_OPTIONS], tot_s[SCG_MAX_OPTIONS], run_s[SCG_MAX_OPTIONS];
    __shared__ uint32_t slot0_s[SCG_MAX_OPTIONS];
    __shared__ uint32_t n_tile_ev;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int mis = (int)(reinterpret_cast<uintptr_t>(ev) & 3);
    const uint32_t *evw = reinterpret_cast<const uint32_t *>(ev - mis);
    const long long nv = (long long)n + mis;            // bytes [mis, nv) of the aligned view are the events
    const long long cta_lo = (long long)blockIdx.x * tiles_per_cta * RING_TILE;
    auto load_word = [&](long long v0) -> uint32_t {    // the aligned word at byte offset v0, bytes outside [mis, nv) cleared
        if (v0 >= nv) return 0u;
        uint32_t w = __ldg(evw + (v0 >> 2));
        if (v0 < mis) w &= 0xFFFFFFFFu << (8 * (mis - (int)v0));
        if (v0 + 4 > nv) w &= 0xFFFFFFFFu >> (8 * (int)(v0 + 4 - nv));
        return w;
    };
    if (tid < SCG_MAX_OPTIONS) { cnt_s[tid] = 0; run_s[tid] = 0; slot0_s[tid] = tid < K ? (uint32_t)((unsigned long long)ex_count[tid] % cap) : 0u; }
    __syncthreads();
    // pass A: this CTA's event count per option
    for (int t = 0; t < tiles_per_cta; ++t) {
        uint32_t w = load_word(cta_lo + (long long)t * RING_TILE + tid * 4) & 0x8F8F8F8Fu;
        while (w & 0x80808080u) {
            const int j = (__ffs(w & 0x80808080u) - 1) >> 3;
            atomicAdd(&cnt_s[(w >> (8 * j)) & SCG_EV_OPT], 1u);
            w &= ~(0xFFu << (8 * j));
        }
    }
    __syncthreads();
    if (tid < SCG_MAX_OPTIONS) scratch[blockIdx.x * SCG_MAX_OPTIONS + tid] = cnt_s[tid];
    // grid barrier (all CTAs are co-resident: gridDim.x <= RING_CTAS): generation-counted arrivals
    __syncthreads();
    if (tid == 0) {
        __threadfence();
        atomicAdd(scratch + RING_BAR, 1u);
        while ((int)(*reinterpret_cast<volatile unsigned int *>(scratch + RING_BAR) - target) < 0) { }
        __threadfence();
    }
    __syncthreads();
    // every CTA's per-option totals: threads (c, o) load them with independent loads, 16 threads form this CTA's bases
    for (int i = tid; i < (int)gridDim.x * SCG_MAX_OPTIONS; i += RING_NT) rk[i] = __ldcg(scratch + i);
    __syncthreads();
    if (tid < SCG_MAX_OPTIONS) {
        uint32_t b = 0, tt = 0;
        for (int c = 0; c < (int)gridDim.x; ++c) {
            const uint32_t v = rk[c * SCG_MAX_OPTIONS + tid];
            if (c < (int)blockIdx.x) b += v;
            tt += v;
        }
        base_s[tid] = b;
        tot_s[tid] = tt;
    }
    __syncthreads();
    // pass B: per tile, the ordered event list (block scan), per-option ranks (warp o walks the list with ballots),
    // then one thread per event places it.  rank = position among this launch's events of the option in (step, env)
    // order; only the last `cap` of them are written, so every slot has exactly one writer.
    for (int t = 0; t < tiles_per_cta; ++t) {
        const long long v0 = cta_lo + (long long)t * RING_TILE + tid * 4;
        const uint32_t w = load_word(v0);
        const uint32_t flags = w & 0x80808080u;
        uint32_t total = 0;
        uint32_t off = ring_block_scan(__popc(flags), warp_sums, &n_tile_ev);
        total = n_tile_ev;
        if (flags) {
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const uint32_t e = (w >> (8 * j)) & 0xFFu;
                if (e & SCG_EV_TERM) lst[off++] = (e & SCG_EV_OPT) | ((e & SCG_EV_HIT) ? 16u : 0u) | ((uint32_t)(tid * 4 + j) << 8);
            }
        }
        __syncthreads();
        if (warp < SCG_MAX_OPTIONS && total) {
            uint32_t run = run_s[warp];
            for (uint32_t i0 = 0; i0 < total; i0 += 32) {
                const uint32_t i = i0 + lane;
                const bool m = i < total && (lst[i] & SCG_EV_OPT) == (uint32_t)warp;
                const unsigned bal = __ballot_sync(0xffffffffu, m);
                if (m) rk[i] = run + __popc(bal & ((1u << lane) - 1u));
                run += __popc(bal);
            }
            if (lane == 0) run_s[warp] = run;
        }
        __syncthreads();
        for (uint32_t i = tid; i < total; i += RING_NT) {
            const uint32_t e = lst[i], o = e & SCG_EV_OPT;
            const uint32_t rank = base_s[o] + rk[i];
            if ((int)o < K && (unsigned long long)rank + cap >= tot_s[o]) {
                const uint32_t slot = (uint32_t)(((unsigned long long)slot0_s[o] + rank) % cap);
                const long long f = cta_lo + (long long)t * RING_TILE + (e >> 8) - mis;     // flat [step][env] index
                const size_t ei = (size_t)o * cap + slot;
                *reinterpret_cast<float2 *>(ex_xy + 2 * ei) = __ldg(pos + f);
                ex_label[ei] = (e & 16u) ? 1 : 0;
            }
        }
        __syncthreads();
    }
    // the last CTA to finish advances the counts (every CTA has read them by then)
    __shared__ bool last;
    if (tid == 0) {
        __threadfence();
        last = atomicInc(scratch + RING_DONE, gridDim.x - 1) == gridDim.x - 1;
    }
    __syncthreads();
    if (last && tid < K) ex_count[tid] += (long long)tot_s[tid];
}

extern "C" int scg_agent_ring(scg_ctx_t *ctx, scg_agent_t *ag, void *stream) {
    if (!ctx || !ag) return SCG_EINVAL;
    if (ag->ring_len < 0 || ag->ring_len > ag->ev_len || ag->ev_len > ag->ev_cap) return SCG_EINVAL;
    const int T = ag->ev_len - ag->ring_len;
    cudaStream_t st = (cudaStream_t)stream;
    if (T > 0 && ag->B > 0) {
        if (!ag->ev_hist || !ag->ev_pos || !ag->ex_xy || !ag->ex_label || !ag->ex_count || ag->example_capacity == 0)
            return SCG_EINVAL;
        if (!ctx->d_ring) {
            SCG_CUDA_OK(cudaMalloc((void **)&ctx->d_ring, RING_WORDS * sizeof(unsigned int)));
            SCG_CUDA_OK(cudaMemsetAsync(ctx->d_ring, 0, RING_WORDS * sizeof(unsigned int), st));
            ctx->ring_gen = 0;
        }
        const long long n = (long long)T * ag->B;
        if (n > 0x7fffffffll) return SCG_ELIMIT;
        const long long n_tiles = (n + 3 + RING_TILE - 1) / RING_TILE;              // (+3: the aligned view may start up to 3 bytes early)
        const int tiles_per_cta = (int)((n_tiles + RING_CTAS - 1) / RING_CTAS);
        const int grid = (int)std::max<long long>(1, (n_tiles + tiles_per_cta - 1) / tiles_per_cta);
        // the barrier counter counts arrivals of all launches so far: this launch is complete at (sum of earlier grids) + grid
        ctx->ring_gen += (unsigned int)grid;
        const size_t off = (size_t)ag->ring_len * ag->B;
        int rcp;
        if ((rcp = scg_prof_push(ctx, 4, st, false))) return rcp;
        k_ring<<<grid, RING_NT, 0, st>>>((int)n, ag->K, ag->ev_hist + off, reinterpret_cast<const float2 *>(ag->ev_pos) + off,
                                         ag->ex_xy, ag->ex_label, reinterpret_cast<long long *>(ag->ex_count),
                                         ag->example_capacity, ctx->d_ring, ctx->ring_gen, tiles_per_cta);
        SCG_LAUNCH_CHECK();
        if ((rcp = scg_prof_push(ctx, 4, st, true))) return rcp;
    }
    ag->ring_len = ag->ev_len;
    // between windows the consumed history is recycled (inside a window its tail still feeds the top-level learner)
    if (ag->win_len == 0) ag->ev_len = ag->ring_len = 0;
    return 0;
}

// ---- the top-level learner's window update --------------------------------------------------------------------------
// Mirrors oracle/option.py OptionSet.top_update.  The step kernel left, for every env-step at which an option terminated
// (event byte), a record (s0, delta_top, option) in win_top.  k_top folds them into the top-level slots of dW:
//     dW[K + o / 5][o % 5][f] += delta_top * phi_f(s0)          and  cnt[K ..] += number of events.
// A CTA scans a contiguous chunk of the event bytes, then processes its events TOP_EB at a time: a few threads build the
// phasor powers exp(i pi c s0_j) of the batch (one sincospi each), then every thread forms its features as products of
// four table entries and accumulates into the CTA's shared-memory copy of the top-level rows; one atomicAdd per touched
// element at the end.
#define TOP_NT 256
#define TOP_EB 8
#define TOP_CTAS 148

template <int N1>
__global__ void __launch_bounds__(TOP_NT) k_top(int n, int K, int top_slots, const uint8_t *__restrict__ ev,
                                                const float4 *__restrict__ top, float *dW, int *cnt) {
    constexpr int F = N1 * N1 * N1 * N1;
    constexpr int FPT = (F + TOP_NT - 1) / TOP_NT;              // features per thread
    extern __shared__ __align__(16) unsigned char smem_raw[];
    float *acc = reinterpret_cast<float *>(smem_raw);           // [K][F]: row j = option j (slot K + j / 5, row j % 5)
    float2 *tab = reinterpret_cast<float2 *>(acc + (((size_t)K * F + 3) & ~(size_t)3));   // [TOP_EB][4][N1], 16-byte aligned
    float *dl = reinterpret_cast<float *>(tab + TOP_EB * 4 * N1);           // [TOP_EB] delta_top
    int *oo = reinterpret_cast<int *>(dl + TOP_EB);                         // [TOP_EB] option
    int *list = oo + TOP_EB;                                                 // [chunk] flat indices of this CTA's events
    __shared__ int n_ev;
    const int tid = threadIdx.x;
    const int chunk = (n + gridDim.x - 1) / gridDim.x;
    const int beg = min(n, (int)blockIdx.x * chunk), end = min(n, beg + chunk);
    for (int i = tid; i < K * F; i += TOP_NT) acc[i] = 0.f;
    if (tid == 0) n_ev = 0;
    __syncthreads();
    for (int i = beg + tid; i < end; i += TOP_NT)
        if (ev[i] & SCG_EV_TERM) list[atomicAdd(&n_ev, 1)] = i;
    __syncthreads();
    const int ne = n_ev;
    if (ne == 0) return;
    // per-thread constants: digits of the owned features, packed 4 x 4 bits
    uint32_t dig[FPT];
#pragma unroll
    for (int k = 0; k < FPT; ++k) {
        int f = tid + k * TOP_NT;
        if (f >= F) f = F - 1;
        const int c3 = f % N1, c2 = (f / N1) % N1, c1 = (f / (N1 * N1)) % N1, c0 = f / (N1 * N1 * N1);
        dig[k] = (uint32_t)c0 | ((uint32_t)c1 << 4) | ((uint32_t)c2 << 8) | ((uint32_t)c3 << 12);
    }
    for (int e0 = 0; e0 < ne; e0 += TOP_EB) {
        const int nb = min(TOP_EB, ne - e0);
        if (tid < nb * 4 * N1) {                                 // table entry (event, dimension j, power c)
            const int e = tid / (4 * N1), r = tid - e * 4 * N1, j = r / N1, c = r - j * N1;
            const float4 s0 = __ldg(top + (size_t)list[e0 + e] * 2);
            const float raw = j == 0 ? s0.x : (j == 1 ? s0.y : (j == 2 ? s0.z : s0.w));
            const float sh = j < 2 ? raw : __fmul_rn(__fadd_rn(raw, 2.0f), 0.25f);
            float sn, cs;
            sincospif((float)c * sh, &sn, &cs);
            tab[tid] = make_float2(cs, sn);
        }
        if (tid < nb) {
            const float4 r1 = __ldg(top + (size_t)list[e0 + tid] * 2 + 1);
            dl[tid] = r1.x;
            oo[tid] = min(max(__float_as_int(r1.y), 0), K - 1);
        }
        __syncthreads();
        for (int e = 0; e < nb; ++e) {
            const float2 *t = tab + e * 4 * N1;
            const float d = dl[e];
            float *row = acc + (size_t)oo[e] * F;
#pragma unroll
            for (int k = 0; k < FPT; ++k) {
                const int f = tid + k * TOP_NT;
                if (f < F) {
                    const float2 p = scg_cmul(scg_cmul(t[dig[k] & 15], t[N1 + ((dig[k] >> 4) & 15)]),
                                              scg_cmul(t[2 * N1 + ((dig[k] >> 8) & 15)], t[3 * N1 + (dig[k] >> 12)]));
                    row[f] = fmaf(d, p.x, row[f]);
                }
            }
        }
        __syncthreads();
    }
    for (int i = tid; i < K * F; i += TOP_NT) {
        const float v = acc[i];
        if (v != 0.f) {
            const int j = i / F, f = i - j * F;
            atomicAdd(dW + ((size_t)(K + j / SCG_A) * SCG_A + (j % SCG_A)) * F + f, v);
        }
    }
    if (tid < top_slots) atomicAdd(cnt + K + tid, ne);
}

// fold the open window's top-level update records (slabs 0 .. win_len-1) into dW / cnt
int scg_launch_top(scg_ctx *ctx, const scg_agent_t *ag, cudaStream_t st) {
    if (ag->top_slots <= 0 || ag->win_len <= 0 || ag->B <= 0) return 0;
    if (!ag->win_top || !ag->ev_hist || ag->ev_len < ag->win_len) return SCG_EINVAL;
    const long long n = (long long)ag->win_len * ag->B;
    if (n > 0x7fffffffll) return SCG_ELIMIT;
    const int grid = (int)std::max<long long>(1, std::min<long long>(TOP_CTAS, (n + 4095) / 4096));
    const int chunk = (int)((n + grid - 1) / grid);
    const int N1 = ctx->order + 1;
    const size_t smem = ((((size_t)ag->K * ctx->F + 3) & ~(size_t)3)) * sizeof(float) + (size_t)TOP_EB * 4 * N1 * sizeof(float2) +
                        TOP_EB * (sizeof(float) + sizeof(int)) + (size_t)chunk * sizeof(int);
    if (smem > 200 * 1024) return SCG_ELIMIT;
    const float4 *top = reinterpret_cast<const float4 *>(ag->win_top);
    int rc = 0;
#define LAUNCH_TOP(N)                                                                                              \
    {                                                                                                               \
        static ScgKernelCfg cfgc = {};                                                                              \
        int per_sm = 0;                                                                                             \
        rc = scg_configure(cfgc, k_top<N>, TOP_NT, smem, &per_sm);                                                  \
        if (!rc) k_top<N><<<grid, TOP_NT, smem, st>>>((int)n, ag->K, ag->top_slots, ag->ev_hist + (size_t)(ag->ev_len - ag->win_len) * ag->B, top, ag->dW, ag->cnt); \
    }
    switch (ctx->order) {
        case 1: LAUNCH_TOP(2); break;
        case 2: LAUNCH_TOP(3); break;
        case 3: LAUNCH_TOP(4); break;
        case 4: LAUNCH_TOP(5); break;
        case 5: LAUNCH_TOP(6); break;
        default: return SCG_ELIMIT;
    }
#undef LAUNCH_TOP
    if (rc) return rc;
    SCG_LAUNCH_CHECK();
    return 0;
}

// ---- the promote-and-fit kernel ---------------------------------------------------------------------------------
#define MANAGE_NT 1024
#define MANAGE_CACHE 8     // examples per thread kept in registers (rings up to 8192 examples never re-read memory)

struct ManageArgs {
    scg_agent_t ag;
    scg_ctl_t *mirror;               // host-mapped copy of ctl, written last
    int world, rank;
    unsigned char *peer[XCHG_MAX_WORLD];
    size_t m_off;
    uint32_t *status;
    long long timeout_cycles;
};

__device__ __forceinline__ void manage_accum(float px, float py, float lab, const float th[SCG_N_PSI], float g[SCG_N_PSI]) {
    const float psi[SCG_N_PSI] = {1.f, px, py, px * px, px * py, py * py};
    float z = 0.f;
#pragma unroll
    for (int j = 0; j < SCG_N_PSI; ++j) z = fmaf(th[j], psi[j], z);
    const float d = 1.0f / (1.0f + expf(-z)) - lab;
#pragma unroll
    for (int j = 0; j < SCG_N_PSI; ++j) g[j] = fmaf(d, psi[j], g[j]);
}

// Sum `nval` (<= XCHG_M_WORDS) per-rank values over the ranks through the peer-memory block: publish, signal every peer,
// wait for every peer, add in rank order (identical on every rank).  vals: shared memory, in/out.  One flag round;
// `seq` is the round counter every rank advances in step.  Returns false (and raises the sticky flag) if a peer never came.
__device__ __forceinline__ bool manage_exchange(const ManageArgs &a, float *vals, int nval, uint32_t &seq, int *s_abort) {
    const int tid = threadIdx.x;
    unsigned char *mine = a.peer[a.rank] + a.m_off;
    seq += 1;
    float *mb = reinterpret_cast<float *>(mine + XCHG_M_BUF) + (seq & 1u) * XCHG_M_WORDS;
    if (tid < nval) mb[tid] = vals[tid];
    __syncthreads();
    if (tid < a.world && tid != a.rank) {
        __threadfence_system();
        st_release_sys(reinterpret_cast<uint32_t *>(a.peer[tid] + a.m_off + XCHG_M_FLAG) + a.rank, seq);
        const uint32_t *lf = reinterpret_cast<const uint32_t *>(mine + XCHG_M_FLAG) + tid;
        const long long t0 = clock64();
        while ((int32_t)(ld_acquire_sys(lf) - seq) < 0) {
            if (clock64() - t0 > a.timeout_cycles) {
                *reinterpret_cast<volatile uint32_t *>(a.status) = 1u;
                __threadfence_system();
                *s_abort = 1;
                break;
            }
        }
    }
    __syncthreads();
    if (tid < nval && !*s_abort) {
        float v = 0.f;
        for (int r = 0; r < a.world; ++r) {
            const float *pb = reinterpret_cast<const float *>(a.peer[r] + a.m_off + XCHG_M_BUF) + (seq & 1u) * XCHG_M_WORDS;
            v += (r == a.rank) ? mb[tid] : ld_sys_f32(pb + tid);
        }
        vals[tid] = v;
    }
    __syncthreads();
    return !*s_abort;
}

__global__ void __launch_bounds__(MANAGE_NT) k_manage(const __grid_constant__ ManageArgs a) {
    __shared__ float sm[33 * 8];
    __shared__ float xv[XCHG_M_WORDS];           // values summed over ranks: a fit step's (6 sums, N) / the merge counts
    __shared__ int s_go, s_abort;
    const scg_agent_t &g = a.ag;
    scg_ctl_t *ctl = g.ctl;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int K = g.K;
    if (tid == 0) {
        const int gi = ctl->n_active;
        int go = 0;
        if (gi < K - 1) {
            // one rank: the live counter; several: the sum over ranks the last exchange left on every GPU
            const long long ns = a.world > 1 ? (long long)g.n_success_global[gi] : (long long)(unsigned int)g.n_success[gi];
            go = ns >= (long long)g.gestation_successes;
        }
        s_go = go;
        s_abort = 0;
    }
    __syncthreads();
    if (s_go) {
        const int gi = ctl->n_active;
        const long long have = g.ex_count[gi];
        const int N = (int)(have < (long long)g.example_capacity ? have : (long long)g.example_capacity);
        const float *X = g.ex_xy + (size_t)gi * g.example_capacity * 2;
        const uint8_t *Y = g.ex_label + (size_t)gi * g.example_capacity;
        float cx[MANAGE_CACHE], cy[MANAGE_CACHE], cl[MANAGE_CACHE];
#pragma unroll
        for (int k = 0; k < MANAGE_CACHE; ++k) {
            const int i = tid + k * MANAGE_NT;
            const bool in = i < N;
            cx[k] = in ? X[2 * i] : 0.f;
            cy[k] = in ? X[2 * i + 1] : 0.f;
            cl[k] = in ? (float)Y[i] : 0.f;
        }
        float th[SCG_N_PSI] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
        unsigned char *mine = a.world > 1 ? a.peer[a.rank] + a.m_off : nullptr;
        uint32_t seq = (a.world > 1) ? *reinterpret_cast<uint32_t *>(mine + XCHG_M_SEQ) : 0u;
        for (int it = 0; it < g.clf_steps && !s_abort; ++it) {
            float gs[SCG_N_PSI] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
#pragma unroll
            for (int k = 0; k < MANAGE_CACHE; ++k)
                if (tid + k * MANAGE_NT < N) manage_accum(cx[k], cy[k], cl[k], th, gs);
            for (int i = tid + MANAGE_CACHE * MANAGE_NT; i < N; i += MANAGE_NT)
                manage_accum(X[2 * i], X[2 * i + 1], (float)Y[i], th, gs);
#pragma unroll
            for (int j = 0; j < SCG_N_PSI; ++j) {
                const float v = scg_warp_sum(gs[j]);
                if (lane == 0) sm[warp * 8 + j] = v;
            }
            __syncthreads();
            if (warp == 0) {
#pragma unroll
                for (int j = 0; j < SCG_N_PSI; ++j) {
                    float v = sm[lane * 8 + j];          // 32 warps
                    v = scg_warp_sum(v);
                    if (lane == 0) xv[j] = v;
                }
                if (lane == 0) xv[6] = (float)N;
            }
            __syncthreads();
            if (a.world > 1) manage_exchange(a, xv, 7, seq, &s_abort);
            const float ntot = xv[6];
            if (ntot > 0.f) {      // no examples anywhere: theta stays 0 (initiation set = everywhere), as in the oracle
#pragma unroll
                for (int j = 0; j < SCG_N_PSI; ++j)
                    th[j] = __fsub_rn(th[j], __fmul_rn(g.clf_lr, __fdiv_rn(xv[j], ntot)));
            }
            __syncthreads();
        }
        if (!s_abort && tid < SCG_N_PSI) g.theta[gi * SCG_N_PSI + tid] = th[tid];
        __syncthreads();
        // the next gestating slot's targets.  chain: {g}.  graph: {g}, every older option whose initiation set holds at
        // least merge_overlap of g's positive examples, and the goal if it lies inside I_g or merge_overlap is 0
        // (oracle/agent.py merged_parents; merge_overlap = 0: all of them)
        uint32_t pm = 1u << gi;
        if (g.graph && !s_abort) {
            if (tid < XCHG_M_WORDS) xv[tid] = 0.f;
            __syncthreads();
            if (g.merge_overlap > 0.f) {
                // inside-counts of the positive examples, per older option (theta_gi was just written: not needed here)
                for (int j = 0; j < gi; ++j) {
                    float c = 0.f, np = 0.f;
                    for (int i = tid; i < N; i += MANAGE_NT) {
                        if (Y[i]) {
                            const float x = X[2 * i], y = X[2 * i + 1];
                            np += 1.f;
                            if (scg_init_logit(g.theta + j * SCG_N_PSI, x, y, __fmul_rn(x, x), __fmul_rn(x, y), __fmul_rn(y, y)) >= 0.f)
                                c += 1.f;
                        }
                    }
                    c = scg_warp_sum(c);
                    np = scg_warp_sum(np);
                    if (lane == 0) { atomicAdd(&xv[j], c); if (j == 0) atomicAdd(&xv[SCG_MAX_OPTIONS], np); }
                }
                if (gi == 0) {      // no older option: still count the positives (not used)
                }
                __syncthreads();
                if (a.world > 1) manage_exchange(a, xv, SCG_MAX_OPTIONS + 1, seq, &s_abort);
                const float npos = xv[SCG_MAX_OPTIONS];
                for (int j = 0; j < gi; ++j) {
                    const float frac = npos > 0.f ? __fdiv_rn(xv[j], npos) : 0.f;
                    if (frac >= g.merge_overlap) pm |= 1u << j;
                }
                const float gx = g.goal_x, gy = g.goal_y;
                const float zg = scg_init_logit(g.theta + gi * SCG_N_PSI, gx, gy, __fmul_rn(gx, gx), __fmul_rn(gx, gy), __fmul_rn(gy, gy));
                // (theta_gi as written above: every thread reads it after the barrier)
                if (zg >= 0.f) pm |= SCG_GOAL_BIT;
            } else {
                pm = ((1u << (gi + 1)) - 1u) | SCG_GOAL_BIT;
            }
        }
        if (a.world > 1 && tid == 0) *reinterpret_cast<uint32_t *>(mine + XCHG_M_SEQ) = seq;
        if (!s_abort && tid == 0) {
            const int n = gi + 1;
            ctl->parents[n] = pm;
            ctl->active_mask |= (1u << gi);
            ctl->n_promotions += 1;
            ctl->last_promotion_step = g.step;
            __threadfence();
            ctl->n_active = n;
        }
    }
    __syncthreads();
    // host mirror: everything but the sequence word, fence, then the sequence word
    if (tid == 0) ctl->manage_calls += 1;
    __syncthreads();
    constexpr int NW = (int)(sizeof(scg_ctl_t) / 4);
    const uint32_t *src = reinterpret_cast<const uint32_t *>(ctl);
    volatile uint32_t *dst = reinterpret_cast<volatile uint32_t *>(a.mirror);
    if (tid < NW && tid != 4) dst[tid] = src[tid];
    __threadfence_system();
    __syncthreads();
    if (tid == 0) dst[4] = src[4];
}

static int ensure_mirror(scg_ctx *ctx) {
    if (ctx->h_ctl) return 0;
    SCG_CUDA_OK(cudaHostAlloc((void **)&ctx->h_ctl, sizeof(scg_ctl_t), cudaHostAllocMapped));
    memset(ctx->h_ctl, 0, sizeof(scg_ctl_t));
    SCG_CUDA_OK(cudaHostGetDevicePointer((void **)&ctx->d_hctl, (void *)ctx->h_ctl, 0));
    return 0;
}

extern "C" int scg_agent_manage(scg_ctx_t *ctx, scg_agent_t *ag, scg_xchg_t *xchg, void *stream) {
    if (!ctx || !ag || !ag->ctl || !ag->theta || !ag->n_success) return SCG_EINVAL;
    if (ag->K + ag->top_slots != ctx->K || ag->K < 1 || ctx->K > SCG_MAX_OPTIONS || ag->clf_steps < 0) return SCG_EINVAL;
    if (xchg && (!ag->n_success_global || xchg->K != ctx->K)) return SCG_EINVAL;
    if (xchg && *xchg->h_status) return SCG_EPEER;
    int rc;
    if ((rc = ensure_mirror(ctx))) return rc;
    if ((rc = scg_agent_ring(ctx, ag, stream))) return rc;    // the rings must hold every example up to this step
    ManageArgs a;
    memset(&a, 0, sizeof(a));
    a.ag = *ag;
    a.mirror = ctx->d_hctl;
    a.world = xchg ? xchg->world : 1;
    a.rank = xchg ? xchg->rank : 0;
    if (xchg) {
        for (int r = 0; r < xchg->world; ++r) {
            if (!xchg->d_peer[r]) return SCG_EINVAL;
            a.peer[r] = xchg->d_peer[r];
        }
        a.m_off = xchg->m_off;
        a.status = xchg->d_status;
        a.timeout_cycles = xchg->timeout_cycles;
    }
    if ((rc = scg_prof_push(ctx, 5, (cudaStream_t)stream, false))) return rc;
    k_manage<<<1, MANAGE_NT, 0, (cudaStream_t)stream>>>(a);
    SCG_LAUNCH_CHECK();
    if ((rc = scg_prof_push(ctx, 5, (cudaStream_t)stream, true))) return rc;
    if (ctx->deterministic) {   // reproducible runs: the host sizes the next launches with exact knowledge
        SCG_CUDA_OK(cudaStreamSynchronize((cudaStream_t)stream));
        ag->n_active = std::max(ag->n_active, (int)ctx->h_ctl->n_active);
    }
    return 0;
}

extern "C" int scg_agent_poll(scg_ctx_t *ctx, scg_ctl_t *out) {
    if (!ctx || !out) return SCG_EINVAL;
    int rc = ensure_mirror(ctx);
    if (rc) return rc;
    const volatile uint32_t *src = reinterpret_cast<const volatile uint32_t *>(ctx->h_ctl);
    uint32_t *dst = reinterpret_cast<uint32_t *>(out);
    // the kernel writes the sequence word last: re-read until it is stable around the copy
    for (int tries = 0; tries < 4; ++tries) {
        const uint32_t s0 = src[4];
        for (size_t i = 0; i < sizeof(scg_ctl_t) / 4; ++i) dst[i] = src[i];
        if (src[4] == s0) break;
    }
    return 0;
}

extern "C" int scg_agent_set_ctl(scg_ctx_t *ctx, scg_agent_t *ag, const scg_ctl_t *in, void *stream) {
    if (!ctx || !ag || !ag->ctl || !in) return SCG_EINVAL;
    if (in->n_active < 0 || in->n_active > ag->K - 1 || (in->active_mask >> ag->K) != 0) return SCG_EINVAL;
    int rc = ensure_mirror(ctx);
    if (rc) return rc;
    // staged through the mirror (pinned): the copy is asynchronous and ordered on the stream
    SCG_CUDA_OK(cudaStreamSynchronize((cudaStream_t)stream));   // a manage kernel in flight may still write the mirror
    memcpy(ctx->h_ctl, in, sizeof(scg_ctl_t));
    SCG_CUDA_OK(cudaMemcpyAsync(ag->ctl, ctx->h_ctl, sizeof(scg_ctl_t), cudaMemcpyHostToDevice, (cudaStream_t)stream));
    SCG_CUDA_OK(cudaStreamSynchronize((cudaStream_t)stream));
    ag->n_active = in->n_active;
    return 0;
}
