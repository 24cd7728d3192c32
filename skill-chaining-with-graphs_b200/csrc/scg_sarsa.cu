// scg_sarsa.cu - K3, the Sarsa(lambda) eligibility-trace sweep and weight-delta accumulation (sm_100a).
//
// Mirrors oracle/option.py OptionSet.update (trace part) and OptionSet.apply; the reference has
// no code (/root/reference/README.md:1-2).  Per env b executing option o with action a:
//     e_b <- gamma*lambda e_b ;  e_b[a] += phi(s_b) ;  dW[o] += delta_b e_b ;  e_b <- 0 if done_b
//
// This is the HBM-bound kernel of the pipeline: the dense trace is A*F fp32 per env, read and
// written once per env-step => 8*A*F (+48 record) algorithmic bytes per env-step.
//
// Design: one CTA sweeps one env at a time; thread t owns the same 16-byte chunk(s) of every env's
// trace, i.e. a fixed set of (action, feature) entries.  Because ownership is exclusive, the CTA's
// dW accumulator (shared memory, [K][A*F]) is updated with plain read-modify-write - no atomics, no
// barriers inside the env loop.  phi(s_b) is rebuilt in registers from the four phasors the
// record carries (only by the threads that own row a), so the only HBM traffic is the trace itself
// plus a 48-byte record per env.  Two envs are in flight per CTA for memory-level parallelism.
// At the end each CTA writes its accumulator to a scratch slab; k_reduce folds the slabs into dW.
#include <stdlib.h>

#include <algorithm>

#include <type_traits>

#include "scg_common.cuh"

template <int VEC> struct VecT;
template <> struct VecT<4> { using type = float4; };
template <> struct VecT<1> { using type = float; };

__device__ __forceinline__ float4 vscale(float4 v, float s) { return make_float4(v.x * s, v.y * s, v.z * s, v.w * s); }
__device__ __forceinline__ float vscale(float v, float s) { return v * s; }
__device__ __forceinline__ float4 vfma(float s, float4 a, float4 c) {
    return make_float4(fmaf(s, a.x, c.x), fmaf(s, a.y, c.y), fmaf(s, a.z, c.z), fmaf(s, a.w, c.w));
}
__device__ __forceinline__ float vfma(float s, float a, float c) { return fmaf(s, a, c); }
__device__ __forceinline__ float4 vzero4() { return make_float4(0.f, 0.f, 0.f, 0.f); }

// z^c for 0 <= c < 8 from (z, z^2, z^4) with selects (no dynamic register indexing)
__device__ __forceinline__ float2 cpow_sel(float2 z1, float2 z2, float2 z4, int c) {
    float2 r = (c & 1) ? z1 : make_float2(1.f, 0.f);
    if (c & 2) r = scg_cmul(r, z2);
    if (c & 4) r = scg_cmul(r, z4);
    return r;
}

struct Phasors {
    float2 z1[4], z2[4], z4[4];
};
__device__ __forceinline__ Phasors make_phasors(float4 r0, float4 r1) {
    Phasors p;
    p.z1[0] = make_float2(r0.x, r0.y); p.z1[1] = make_float2(r0.z, r0.w);
    p.z1[2] = make_float2(r1.x, r1.y); p.z1[3] = make_float2(r1.z, r1.w);
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        p.z2[j] = scg_cmul(p.z1[j], p.z1[j]);
        p.z4[j] = scg_cmul(p.z2[j], p.z2[j]);
    }
    return p;
}
// phi for the feature whose base-N1 digits are packed 4 bits each (c0 lowest nibble)
__device__ __forceinline__ float phi_digits(const Phasors &p, uint32_t d) {
    float2 a = cpow_sel(p.z1[0], p.z2[0], p.z4[0], d & 15);
    float2 b = cpow_sel(p.z1[1], p.z2[1], p.z4[1], (d >> 4) & 15);
    float2 c = cpow_sel(p.z1[2], p.z2[2], p.z4[2], (d >> 8) & 15);
    float2 e = cpow_sel(p.z1[3], p.z2[3], p.z4[3], (d >> 12) & 15);
    float2 ab = scg_cmul(a, b), ce = scg_cmul(c, e);
    return fmaf(ab.x, ce.x, -ab.y * ce.y);
}
template <int N1>
__device__ __forceinline__ uint32_t pack_digits(int f) {
    int c3 = f % N1; f /= N1;
    int c2 = f % N1; f /= N1;
    int c1 = f % N1; f /= N1;
    return (uint32_t)f | ((uint32_t)c1 << 4) | ((uint32_t)c2 << 8) | ((uint32_t)c3 << 12);
}

__device__ __forceinline__ float4 add_phi(float4 e, const Phasors &p, const uint32_t d[4]) {
    e.x += phi_digits(p, d[0]); e.y += phi_digits(p, d[1]);
    e.z += phi_digits(p, d[2]); e.w += phi_digits(p, d[3]);
    return e;
}
__device__ __forceinline__ float add_phi(float e, const Phasors &p, const uint32_t d[1]) { return e + phi_digits(p, d[0]); }

template <int N1, int VEC, int CPT, int NT, int U>
__global__ void __launch_bounds__(NT) k_trace(int B, int K, const float4 *__restrict__ rec, float *trace,
                                              float *__restrict__ partial, float gl) {
    using V = typename VecT<VEC>::type;
    constexpr int F = N1 * N1 * N1 * N1;
    constexpr int AF = SCG_A * F;
    constexpr int NCH = AF / VEC;
    static_assert(AF % VEC == 0, "vector width must divide the trace");
    static_assert(VEC == 1 || F % VEC == 0, "a vector must not straddle action rows");
    static_assert(NT * CPT >= NCH, "not enough threads");
    extern __shared__ __align__(128) unsigned char smem_raw[];
    V *acc = reinterpret_cast<V *>(smem_raw);  // [K][NCH]

    // per-thread constants: which (row, features) each owned chunk covers
    int row[CPT];
    uint32_t dig[CPT][VEC];
    bool own[CPT];
#pragma unroll
    for (int i = 0; i < CPT; ++i) {
        int c = threadIdx.x + i * NT;
        own[i] = c < NCH;
        int e0 = (own[i] ? c : 0) * VEC;
        row[i] = e0 / F;
#pragma unroll
        for (int v = 0; v < VEC; ++v) dig[i][v] = pack_digits<N1>((e0 + v) % F);
    }
    for (int i = threadIdx.x; i < K * NCH; i += NT) {
        if constexpr (VEC == 4) acc[i] = vzero4(); else acc[i] = 0.f;
    }
    __syncthreads();

    const int stride = gridDim.x;
    for (int b0 = blockIdx.x; b0 < B; b0 += U * stride) {
        // issue every load of the U envs in flight before touching any of them
        float2 dm[U];          // (delta, meta) of each env: the third 16-byte piece of its record
        V e[U][CPT];
        bool live[U];
#pragma unroll
        for (int u = 0; u < U; ++u) {
            int b = b0 + u * stride;
            live[u] = b < B;
            if (live[u]) {
                dm[u] = __ldg(reinterpret_cast<const float2 *>(rec + (size_t)b * 3 + 2));
                const V *tp = reinterpret_cast<const V *>(trace + (size_t)b * AF);
#pragma unroll
                for (int i = 0; i < CPT; ++i)
                    if (own[i]) e[u][i] = tp[threadIdx.x + i * NT];
            }
        }
#pragma unroll
        for (int u = 0; u < U; ++u) {
            if (!live[u]) continue;
            int b = b0 + u * stride;
            uint32_t meta = __float_as_uint(dm[u].y);
            if (!(meta & SCG_META_ACTIVE)) continue;
            float delta = dm[u].x;
            int a = meta & 7, o = (meta >> 8) & 0xFF;
            bool zero_after = (meta & SCG_META_ZERO_AFTER) != 0;
            V *tp = reinterpret_cast<V *>(trace + (size_t)b * AF);
            V *ap = acc + (size_t)o * NCH;
#pragma unroll
            for (int i = 0; i < CPT; ++i) {
                if (!own[i]) continue;
                int c = threadIdx.x + i * NT;
                V v = vscale(e[u][i], gl);
                if (row[i] == a) {   // only the owners of row a rebuild phi(s_b), from the record's phasors
                    Phasors ph = make_phasors(__ldg(rec + (size_t)b * 3), __ldg(rec + (size_t)b * 3 + 1));
                    v = add_phi(v, ph, dig[i]);
                }
                ap[c] = vfma(delta, v, ap[c]);
                if (zero_after) {
                    if constexpr (VEC == 4) v = vzero4(); else v = 0.f;
                }
                tp[c] = v;
            }
        }
    }
    __syncthreads();
    V *out = reinterpret_cast<V *>(partial + (size_t)blockIdx.x * K * AF);
    for (int i = threadIdx.x; i < K * NCH; i += NT) out[i] = acc[i];
}

// ---- windowed (forward-view) sweep: the agent pipeline's K3 --------------------------------------
// Mirrors oracle/option.py OptionSet.flush.  The fused step kernel left, for each of the T steps of the
// window, a 32-byte record per env (state, delta, action / option / termination bits).  One CTA sweeps
// one env at a time; thread t owns the same VEC features of all A rows of every env's trace (registers)
// and of the CTA's dW accumulator (shared memory, [K][A*F], plain read-modify-write: ownership is
// exclusive).  Work items are (env, block of 8 steps); per item:
//   build  the pair tables P01[t][c0][c1] = exp(i pi (c0 s0 + c1 s1)) and P23[t][c2][c3] = exp(i pi (c2 s2 + c3 s3))
//          of the NEXT item, one SFU sincos per entry; for an env's first block a scan warp also runs the backward
//          recursion G_t = delta_t + (done_t ? 0 : gl G_{t+1}), the trace coefficients c_t and (windows of <= 8
//          steps) sorts the contributing steps by (segment, action) into a list
//   main   phi_f = Re(P01[f / N1^2] P23[f % N1^2]);  d[a_t][f] += G_t phi_f;  e[a_t][f] += c_t phi_f
//          with d (the env's dW contribution, seeded with the carry-in gl G_0 e_start) and e in registers;
//          d is added to the shared accumulator once per env (or when the option changes)
// The tables are double-buffered and the records of the following items are prefetched into registers, so there is one
// barrier per item and no phase waits on a load it has just issued.  The next trace is prefetched either into L2 by one
// bulk-prefetch instruction (L2PF: order 3, where the 20 registers it frees buy two more CTAs per SM) or into registers
// (order 5: one CTA per SM, nothing else would cover an L2 hit).  With two warps per env (order 3, SPLIT) warp 0 runs
// the scan and warp 1 builds all the table entries, so neither waits for the other at the barrier; with many warps per
// env (orders 4, 5) an extra warp does nothing but the scan (CTRL).  A feature costs two 8-byte shared loads and two
// FP32 ops to form, and the dense trace crosses HBM once per window: 8*A*F/T + 32 algorithmic bytes per env-step.
#define SCG_WIN_TB 8   // steps per table block
#ifndef SCG_WIN_PF
#define SCG_WIN_PF 1
#endif

struct WinItem {   // a work item of the sweep: block `blk` of 8 steps of env `b`
    int b, blk;
};

// per-env control block written by the scan warp, read by every thread in the main phase (double-buffered)
template <int NS>                      // NS: steps a window can have (8 for the single-block sweep, SCG_WIN_MAX otherwise)
struct WinCtlT {
    float2 gc[NS];                     // (G_t, c_t) per step
    uint32_t mask[NS][8];              // [segment][action]: steps of that segment taken with that action (5 used)
    uint32_t seg_o[NS];                // option of each segment (an env changes option only after a termination)
    float scale, carry;                // e <- scale * e_start;  d <- carry * e_start
    int nseg, pad;
};

// Single-block sweep (T <= 8): the scan warp sorts the window's steps by (segment, action) into a list, so that the main
// phase walks plain index ranges with static register rows - one broadcast 16-byte load per step carries G_t, c_t and
// the byte offset of the step's tables, and there is no bit-mask bookkeeping in the loop.
struct WinCtlList {
    float4 ent[SCG_WIN_TB + 1];        // (G_t, c_t, byte offset of step t's pair tables, -) in (segment, action, t) order
    uint8_t rng[SCG_WIN_TB][8];        // [segment][action]: first list index of that action's steps; [5] = end
    uint32_t seg_o[SCG_WIN_TB];
    float scale, carry;
    int nseg, pad;
};

template <int N1, int VEC, int NT, int CH, bool MULTI, int MINB, bool CTRL, bool LISTP, bool L2PF, bool SPLITP>
__global__ void __launch_bounds__(NT + (CTRL ? 32 : 0), MINB) k_window(int B, int K, int T, const float4 *__restrict__ rec, float *trace,
                                               float *__restrict__ partial, float gl, float *dW, int K_all) {
    using V = typename VecT<VEC>::type;
    constexpr bool LIST = LISTP && !MULTI;
    using WinCtl = std::conditional_t<LIST, WinCtlList, WinCtlT<MULTI ? SCG_WIN_MAX : SCG_WIN_TB>>;
    constexpr int F = N1 * N1 * N1 * N1;
    constexpr int AF = SCG_A * F;
    constexpr int NCHR = F / VEC;          // chunks per action row
    constexpr int NN = N1 * N1;
    constexpr int TABN = SCG_WIN_TB * 2 * NN;             // entries of one table buffer
    // Storage of one step's tables, in float2: P01 (NN entries) then P23 in two halves of 16-byte groups (entries
    // (4g, 4g+1) of every group g of four, then entries (4g+2, 4g+3)).  DUP9 (order 5): the 9 groups of a half cannot
    // sit in 8 distinct bank quads, and a quarter-warp reads 8 consecutive groups (mod 9) with one 16-byte load: 7 of 9
    // quarters hit groups 0 and 8 together - two wavefronts (6.4 per load measured).  So a half is 16 slots: groups 0..7,
    // then eight copies of group 8, copy m in the bank quad of group m; the thread that needs group 8 reads the copy in
    // the quad its quarter leaves free.
    constexpr bool DUP9 = VEC == 4 && NN / 4 == 9;
    constexpr int HS = DUP9 ? 32 : NN / 2;                // float2 per half of P23
    constexpr int TSZ = VEC == 4 ? NN + 2 * HS : 2 * NN;  // float2 per step
    constexpr int TABS = SCG_WIN_TB * TSZ;                // float2 per table buffer
    // SPLIT (two warps per env): warp 0 runs the per-env scan, warp 1 builds all the table entries - the two jobs take
    // about as long, so neither warp waits for the other at the per-item barrier.  The 16 state pairs the entries need
    // (8 steps x positions / velocities) are loaded by 16 lanes, one each, and handed round with shuffles.
    constexpr bool SPLIT = SPLITP && !CTRL && !MULTI && NT == 64 && 2 * NN == 32;
    constexpr int BT = SPLIT ? 32 : NT;                   // threads that build table entries
    constexpr int EPT = (TABN + BT - 1) / BT;             // table entries built per builder thread
    constexpr int TSTRIDE = TSZ * (int)sizeof(float2);    // bytes between the tables of consecutive steps
    static_assert(F % VEC == 0 && NT * CH >= NCHR && NT >= 32 && SCG_WIN_MAX <= 32, "layout");
    static_assert(VEC == 1 || NN % VEC == 0, "a chunk shares its (c0, c1) digits");
    extern __shared__ __align__(128) unsigned char smem_raw[];
    float *acc = reinterpret_cast<float *>(smem_raw);                                      // [K][AF]
    float2 *tab = reinterpret_cast<float2 *>(acc + (((size_t)K * AF + 3) & ~(size_t)3));   // [2][TABS]
    WinCtl *ctl = reinterpret_cast<WinCtl *>(tab + 2 * TABS);                              // [2]
    float *glpow = reinterpret_cast<float *>(ctl + 2);                                     // [36]: gl^n

    const int tid = threadIdx.x;
    // CTRL: one extra warp (threads NT .. NT+31) only runs the per-env scan, so that no work warp is slower than the
    // others at the per-env barrier (worth it when a CTA has many work warps: orders 4 and 5)
    constexpr int NTT = NT + (CTRL ? 32 : 0);
    const bool scan_warp = CTRL ? (tid >= NT) : (tid < 32);
    const bool worker = tid < NT;
    const bool builder = SPLIT ? tid >= 32 : worker;
    const int bt = SPLIT ? tid - 32 : tid;                // builder index (SPLIT: lane = half * NN + ij)
    // thread t owns chunks t, t + NT, ... of every action row (each a coalesced access across the CTA)
    bool own[CH];
#pragma unroll
    for (int j = 0; j < CH; ++j) own[j] = tid + j * NT < NCHR;
    const int NBLK = MULTI ? (T + SCG_WIN_TB - 1) / SCG_WIN_TB : 1;
    // per-thread constants: table offsets (bytes) of the owned features' (c0, c1) and (c2, c3) entries ...
    int off01[CH], off23[CH];
#pragma unroll
    for (int j = 0; j < CH; ++j) {
        const int f0 = (own[j] ? tid + j * NT : 0) * VEC;
        off01[j] = (f0 / NN) * (int)sizeof(float2);
        // P23: scalar form one float2 per (c2, c3); vector form: the 16-byte slot of the chunk's group in the first half
        // (the second half is HS float2 further on); DUP9: group 8 is read from the copy in the bank quad that the
        // other seven lanes of this quarter-warp leave free
        if constexpr (VEC == 4) {
            const int c = own[j] ? tid + j * NT : 0;
            int slot = c % (NN / 4);
            if (DUP9 && slot == 8) slot = 8 + (8 * (c / 8) + 8) % 9;
            off23[j] = NN * (int)sizeof(float2) + slot * 16;
        } else {
            off23[j] = (NN + f0 % NN) * (int)sizeof(float2);
        }
    }
    // ... and, for the EPT table entries this thread builds: step within the block, which state pair, digits,
    // the affine map of the raw state pair to [0, 1] (positions: identity; velocities: (v + 2) / 4) folded into the
    // digit coefficients (x = ca s0 + cb s1 + cc on the raw pair), and where the entry goes.
    // REG: the threads of a CTA tile the steps' tables exactly (NT a multiple of the 2 NN entries of a step), so all of
    // this is the same for every entry a thread builds, up to a constant step stride - scalars instead of arrays.
    constexpr bool REG = BT % (2 * NN) == 0 && TABN % BT == 0 && !DUP9;
    constexpr int EN = REG ? 1 : EPT;
    int e_tt[EN], e_half[EN], e_at[EN];
    bool e_dup[EN];                                      // DUP9: an entry of group 8 - stored eight times
    float e_ca[EN], e_cb[EN], e_cc[EN];
#pragma unroll
    for (int k = 0; k < EN; ++k) {
        int idx = (builder ? bt : 0) + k * BT;
        if (idx >= TABN) idx = TABN - 1;                 // duplicate work on the last entry, never out of range
        const int tt = idx / (2 * NN), rem = idx - tt * 2 * NN, half = rem / NN, ij = rem - half * NN;
        e_tt[k] = tt; e_half[k] = half;
        const float mul = half ? 0.25f : 1.f, add = half ? 0.5f : 0.f;
        e_ca[k] = (float)(ij / N1) * mul; e_cb[k] = (float)(ij % N1) * mul;
        e_cc[k] = (float)(ij / N1 + ij % N1) * add;
        int at = tt * TSZ + rem;
        if (VEC == 4 && half) {
            const int g = ij >> 2, kk = ij & 3;
            at = tt * TSZ + NN + (kk >> 1) * HS + 2 * g + (kk & 1);
        }
        e_at[k] = at;
        e_dup[k] = DUP9 && half && (ij >> 2) == 8;
    }
    auto ent_tt = [&](int k) { return REG ? e_tt[0] + k * (BT / (2 * NN)) : e_tt[REG ? 0 : k]; };
    auto ent_at = [&](int k) { return REG ? e_at[0] + k * BT : e_at[REG ? 0 : k]; };
    for (int i = tid; i < K * AF; i += NTT) acc[i] = 0.f;
    if (tid == 0) {
        float p = 1.f;
        for (int n = 0; n < 36; ++n) { glpow[n] = p; p *= gl; }
    }

    const int stride = gridDim.x;
    auto advance = [&](WinItem it) {
        if constexpr (MULTI) {
            if (++it.blk == NBLK) { it.blk = 0; it.b += stride; }
        } else {
            it.b += stride;
        }
        return it;
    };

    // prefetch registers
    float2 r_sv[SPLIT ? 1 : EPT];                        // state pair of the steps this thread builds entries for
    float2 r_dm = make_float2(0.f, 0.f);                 // (delta, meta) of step tid (first block of an env only)
    V e[CH][SCG_A], d[CH][SCG_A];
    V e_nx[L2PF ? 1 : CH][L2PF ? 1 : SCG_A];            // register prefetch of the next env's trace (!L2PF)
    uint32_t o_cur = 0;
#pragma unroll
    for (int j = 0; j < CH; ++j) {
#pragma unroll
        for (int r = 0; r < SCG_A; ++r) {
            if constexpr (VEC == 4) { e[j][r] = vzero4(); d[j][r] = vzero4(); }
            else { e[j][r] = 0.f; d[j][r] = 0.f; }
        }
    }

    auto load_rec = [&](WinItem it) {                    // records of a work item -> registers
        if constexpr (SPLIT) {
            if (tid >= 32 && tid < 32 + 2 * SCG_WIN_TB) {                 // lane l: step l / 2, pair l & 1
                const int t = min((tid - 32) >> 1, T - 1);
                r_sv[0] = __ldg(reinterpret_cast<const float2 *>(rec + ((size_t)t * B + it.b) * 2) + (tid & 1));
            }
        } else if (worker) {
#pragma unroll
            for (int k = 0; k < EPT; ++k) {
                const int t = min(it.blk * SCG_WIN_TB + ent_tt(k), T - 1);
                r_sv[k] = __ldg(reinterpret_cast<const float2 *>(rec + ((size_t)t * B + it.b) * 2) + e_half[REG ? 0 : k]);
            }
        }
        if (it.blk == 0 && scan_warp && (tid & 31) < T)
            r_dm = __ldg(reinterpret_cast<const float2 *>(rec + ((size_t)(tid & 31) * B + it.b) * 2 + 1));
    };
    // The trace of an env is pulled into L2 two items ahead by one bulk-prefetch instruction (no registers, no shared
    // memory); the item itself then loads it straight into the trace registers, an L2 hit whose latency the table build
    // covers.  (Holding the next trace in registers instead cost 20 registers per thread - one CTA per SM at order 3.)
    auto prefetch_trace = [&](int b) {
        constexpr int BYTES = AF * (int)sizeof(float);
        if constexpr (BYTES % 16 == 0) {
            if (tid == NT - 1)
                asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(trace + (size_t)b * AF), "n"(BYTES)
                             : "memory");
        } else {                                         // odd N1: the env's trace is only 4-byte aligned - line by line
            const char *base = reinterpret_cast<const char *>(trace + (size_t)b * AF);
            for (int ofs = tid * 128; ofs < BYTES; ofs += NT * 128)
                asm volatile("prefetch.global.L2 [%0];" ::"l"(base + ofs) : "memory");
        }
    };
    auto load_trace = [&](int b) {
        const V *tp = reinterpret_cast<const V *>(trace + (size_t)b * AF);
#pragma unroll
        for (int j = 0; j < CH; ++j) {
            if (own[j]) {
#pragma unroll
                for (int r = 0; r < SCG_A; ++r) {
                    if constexpr (L2PF) e[j][r] = tp[r * NCHR + tid + j * NT];
                    else e_nx[j][r] = tp[r * NCHR + tid + j * NT];
                }
            }
        }
    };
    // registers -> pair tables (buffer `par`) and, for an env's first block, its control block (buffer `epar`)
    auto build = [&](WinItem it, int par, int epar) {
        float2 *tb = tab + par * TABS;
        if (builder) {
#pragma unroll
            for (int k = 0; k < EPT; ++k) {
                // exp(i pi x): exact reduction of x to [-1, 1], then the SFU (abs error ~4e-7)
                const int kc = REG ? 0 : k;
                float2 sv;
                if constexpr (SPLIT) {                   // entry k of this lane is step k: its pair sits in lane 2k + half
                    sv.x = __shfl_sync(0xffffffffu, r_sv[0].x, 2 * k + e_half[0]);
                    sv.y = __shfl_sync(0xffffffffu, r_sv[0].y, 2 * k + e_half[0]);
                } else {
                    sv = r_sv[k];
                }
                const float x = fmaf(e_ca[kc], sv.x, fmaf(e_cb[kc], sv.y, e_cc[kc]));
                const float xr = 3.14159265358979f * fmaf(-2.f, rintf(0.5f * x), x);
                if (REG || bt + k * BT < TABN) {
                    const float2 v = make_float2(__cosf(xr), __sinf(xr));
                    tb[ent_at(k)] = v;
                    if constexpr (DUP9) {
                        if (e_dup[kc]) {
#pragma unroll
                            for (int mq = 1; mq < 8; ++mq) tb[ent_at(k) + 2 * mq] = v;
                        }
                    }
                }
            }
        }
        if (it.blk == 0 && scan_warp) {
            // warp 0, lane = step: backward recursion G_t = delta_t + m_t G_{t+1} as a suffix scan over (m, delta)
            // pairs, m_t = 0 after a termination, gl otherwise (1 for a step the env sat out)
            WinCtl &cb = ctl[epar];
            const unsigned FULL = 0xffffffffu;
            const int t = tid & 31;
            const uint32_t meta = (t < T) ? __float_as_uint(r_dm.y) : 0u;
            const bool act = (meta & SCG_META_ACTIVE) != 0;
            const bool dn = act && (meta & SCG_META_ZERO_AFTER);
            const uint32_t a = meta & 7u, o = (meta >> 8) & 0xFFu;
            float D = act ? r_dm.x : 0.f, M = !act ? 1.f : (dn ? 0.f : gl);
#pragma unroll
            for (int off = 1; off < (MULTI ? 32 : SCG_WIN_TB); off <<= 1) {   // T <= 8 steps without MULTI: 3 rounds
                const float D2 = __shfl_down_sync(FULL, D, off), M2 = __shfl_down_sync(FULL, M, off);
                if (t + off < 32) { D = fmaf(M, D2, D); M *= M2; }
            }
            const unsigned am = __ballot_sync(FULL, act), dm = __ballot_sync(FULL, dn);
            const bool dead = (dm >> t) != 0;                     // a termination at this step or later in the window
            const float cf = (act && !dead) ? glpow[__popc((am >> t) >> 1)] : 0.f;
            const float G = act ? D : 0.f;
            // segments: maximal runs of steps under the same option
            const unsigned before = am & ((1u << t) - 1u);
            const int prev = before ? 31 - __clz(before) : t;
            const uint32_t o_prev = __shfl_sync(FULL, o, prev);
            const unsigned chg = __ballot_sync(FULL, act && before && o != o_prev);
            const int seg = __popc(chg & ((2u << t) - 1u));
            const int nseg = am ? __popc(chg) + 1 : 0;
            const bool nz = act && (G != 0.f || cf != 0.f);
            if constexpr (!LIST) {
                if (t < T) cb.gc[t] = make_float2(G, cf);
                for (int sgm = 0; sgm < nseg; ++sgm) {
                    const unsigned in_seg = __ballot_sync(FULL, act && seg == sgm);
                    const uint32_t os = __shfl_sync(FULL, o, __ffs(in_seg) - 1);
                    unsigned mine = 0;
#pragma unroll
                    for (int r = 0; r < SCG_A; ++r) {
                        const unsigned mr = __ballot_sync(FULL, nz && seg == sgm && a == (uint32_t)r);
                        if (t == r) mine = mr;
                    }
                    if (t < SCG_A) cb.mask[sgm][t] = mine;
                    if (t == 0) cb.seg_o[sgm] = os;
                }
            } else {
                // counting sort of the contributing steps by (segment, action, t): segments are runs in time
                uint32_t m[SCG_A], nzm = 0;
#pragma unroll
                for (int r = 0; r < SCG_A; ++r) { m[r] = __ballot_sync(FULL, nz && a == (uint32_t)r); nzm |= m[r]; }
                const unsigned below = (1u << t) - 1u;
                for (int sgm = 0; sgm < nseg; ++sgm) {
                    const unsigned in_seg = __ballot_sync(FULL, act && seg == sgm);
                    const int first = __ffs(in_seg) - 1;
                    const uint32_t os = __shfl_sync(FULL, o, first);
                    const int base = __popc(nzm & ((1u << first) - 1u));
                    int c_lane = base, c_own = base;          // list index where action `t` / this step's action starts
                    uint32_t m_own = 0;
#pragma unroll
                    for (int r = 0; r < SCG_A; ++r) {
                        const int pc = __popc(m[r] & in_seg);
                        if (r < t) c_lane += pc;
                        if ((uint32_t)r < a) c_own += pc;
                        if ((uint32_t)r == a) m_own = m[r];
                    }
                    if (t <= SCG_A) cb.rng[sgm][t] = (uint8_t)c_lane;
                    if (nz && seg == sgm)
                        cb.ent[c_own + __popc(m_own & in_seg & below)] =
                            make_float4(G, cf, __int_as_float(par * TABS * (int)sizeof(float2) + t * TSTRIDE), 0.f);
                    if (t == 0) cb.seg_o[sgm] = os;
                }
            }
            if (t == 0) {
                cb.scale = dm ? 0.f : glpow[__popc(am)];
                cb.carry = am ? gl * D : 0.f;
                cb.nseg = nseg;
            }
        }
    };
    auto flush_d = [&]() {
        if (o_cur >= (uint32_t)K) {
            // an option promoted after the host sized this launch (the accumulator covers options 0 .. K-1 only): its
            // contribution goes straight to dW with atomics (rare and transient; ids beyond K_all are ignored)
            if (o_cur < (uint32_t)K_all) {
                float *gp = dW + (size_t)o_cur * AF;
#pragma unroll
                for (int j = 0; j < CH; ++j) {
                    if (own[j]) {
#pragma unroll
                        for (int r = 0; r < SCG_A; ++r) {
                            float *q = gp + (size_t)(r * NCHR + tid + j * NT) * VEC;
                            if constexpr (VEC == 4) {
                                atomicAdd(q, d[j][r].x); atomicAdd(q + 1, d[j][r].y);
                                atomicAdd(q + 2, d[j][r].z); atomicAdd(q + 3, d[j][r].w);
                            } else {
                                atomicAdd(q, d[j][r]);
                            }
                        }
                    }
                }
            }
            return;
        }
        V *ap = reinterpret_cast<V *>(acc + (size_t)o_cur * AF);
#pragma unroll
        for (int j = 0; j < CH; ++j) {
            if (own[j]) {
#pragma unroll
                for (int r = 0; r < SCG_A; ++r) {
                    V cur = ap[r * NCHR + tid + j * NT];
                    if constexpr (VEC == 4)
                        cur = make_float4(cur.x + d[j][r].x, cur.y + d[j][r].y, cur.z + d[j][r].z, cur.w + d[j][r].w);
                    else cur = cur + d[j][r];
                    ap[r * NCHR + tid + j * NT] = cur;
                }
            }
        }
    };

    WinItem cur = {(int)blockIdx.x, 0};
    if (cur.b < B) {
        load_rec(cur);
        if constexpr (!L2PF) load_trace(cur.b);
    }
    __syncthreads();                                     // accumulator zeroed, glpow ready
    WinItem nxt = advance(cur);
    if (L2PF && nxt.b < B) prefetch_trace(nxt.b);
    if (cur.b < B) build(cur, 0, 0);
    if (nxt.b < B) load_rec(nxt);
    __syncthreads();

    int par = 0, epar = 0;
    while (cur.b < B) {
        const WinItem nx2 = advance(nxt);
        if constexpr (L2PF) {
            if (cur.blk == 0) load_trace(cur.b);
            if (nx2.blk == 0 && nx2.b < B) prefetch_trace(nx2.b);
        } else if (cur.blk == 0) {
#pragma unroll
            for (int j = 0; j < CH; ++j) {
#pragma unroll
                for (int r = 0; r < SCG_A; ++r) e[j][r] = e_nx[j][r];
            }
        }
        // next item's tables (into the other buffer), then the loads for the items after it
        if (nxt.b < B) build(nxt, par ^ 1, nxt.blk == 0 ? epar ^ 1 : epar);
        if (nx2.b < B) load_rec(nx2);
        if constexpr (!L2PF) {
            if (nxt.blk == 0 && nxt.b < B) load_trace(nxt.b);
        }
        // ---- main ----
        const WinCtl &cb = ctl[epar];
        const char *tb0 = reinterpret_cast<const char *>(tab + par * TABS);
        [[maybe_unused]] const float2 *gc = nullptr;
        if constexpr (!LIST) gc = cb.gc + cur.blk * SCG_WIN_TB;
        if (cur.blk == 0) {
            const float carry = cb.carry, scale = cb.scale;
            o_cur = cb.seg_o[0];
#pragma unroll
            for (int j = 0; j < CH; ++j) {
#pragma unroll
                for (int r = 0; r < SCG_A; ++r) {
                    d[j][r] = vscale(e[j][r], carry);    // carry-in: gl G_0 e_start
                    e[j][r] = vscale(e[j][r], scale);
                }
            }
        }
        const int nseg = cb.nseg;
        auto change_option = [&](uint32_t so) {          // the env changed option inside the window (after a termination)
            flush_d();
#pragma unroll
            for (int j = 0; j < CH; ++j) {
#pragma unroll
                for (int r = 0; r < SCG_A; ++r) {
                    if constexpr (VEC == 4) d[j][r] = vzero4(); else d[j][r] = 0.f;
                }
            }
            o_cur = so;
        };
        // one step's contribution to row r of the owned chunks; tb: the step's pair tables
        auto step_row = [&](const char *tb, float G, float c, V (&dd)[CH][SCG_A], V (&ee)[CH][SCG_A], int r) {
#pragma unroll
            for (int j = 0; j < CH; ++j) {               // independent chunks: their loads and FMAs interleave
                const float2 p = *reinterpret_cast<const float2 *>(tb + off01[j]);
                V ph;
                if constexpr (VEC == 4) {
                    const float4 q01 = *reinterpret_cast<const float4 *>(tb + off23[j]);
                    const float4 q23 = *reinterpret_cast<const float4 *>(tb + off23[j] + HS * (int)sizeof(float2));
                    ph = make_float4(fmaf(p.x, q01.x, -p.y * q01.y), fmaf(p.x, q01.z, -p.y * q01.w),
                                     fmaf(p.x, q23.x, -p.y * q23.y), fmaf(p.x, q23.z, -p.y * q23.w));
                } else {
                    const float2 q = *reinterpret_cast<const float2 *>(tb + off23[j]);
                    ph = fmaf(p.x, q.x, -p.y * q.y);
                }
                dd[j][r] = vfma(G, ph, dd[j][r]);
                ee[j][r] = vfma(c, ph, ee[j][r]);
            }
        };
        if constexpr (LIST) {
            const char *tbase = reinterpret_cast<const char *>(tab);
            for (int sgm = 0; sgm < nseg; ++sgm) {
                const uint2 rr = *reinterpret_cast<const uint2 *>(cb.rng[sgm]);
                int lo = rr.x & 0xFF;
                if (lo == (int)((rr.y >> 8) & 0xFF)) continue;         // no contributing step in this segment
                const uint32_t so = cb.seg_o[sgm];
                if (so != o_cur) change_option(so);
#if SCG_WIN_PF
                float4 en = cb.ent[lo];                  // the list is contiguous across the actions: always one ahead
#endif
#pragma unroll
                for (int r = 0; r < SCG_A; ++r) {        // static register rows: no dynamic indexing, no switch
                    const int hi = r < 3 ? (rr.x >> (8 * (r + 1))) & 0xFF : (rr.y >> (8 * (r - 3))) & 0xFF;
#pragma unroll 1
                    for (int i = lo; i < hi; ++i) {      // (two steps per trip, to spare the copy: measured equal)
#if SCG_WIN_PF
                        const float4 nx = cb.ent[i + 1];
                        step_row(tbase + __float_as_int(en.z), en.x, en.y, d, e, r);
                        en = nx;
#else
                        const float4 en = cb.ent[i];
                        step_row(tbase + __float_as_int(en.z), en.x, en.y, d, e, r);
#endif
                    }
                    lo = hi;
                }
            }
        } else {
        for (int sgm = 0; sgm < nseg; ++sgm) {
            // this segment's steps inside this block, one bit mask per action
            uint32_t mk[SCG_A];
            {
                const uint4 m4 = *reinterpret_cast<const uint4 *>(cb.mask[sgm]);
                mk[0] = m4.x; mk[1] = m4.y; mk[2] = m4.z; mk[3] = m4.w; mk[4] = cb.mask[sgm][4];
            }
            uint32_t any = 0;
#pragma unroll
            for (int r = 0; r < SCG_A; ++r) {
                if constexpr (MULTI) mk[r] = (mk[r] >> (cur.blk * SCG_WIN_TB)) & ((1u << SCG_WIN_TB) - 1u);
                any |= mk[r];
            }
            if (!any) continue;
            const uint32_t so = cb.seg_o[sgm];
            if (so != o_cur) change_option(so);
#pragma unroll
            for (int r = 0; r < SCG_A; ++r) {
                uint32_t mm = mk[r];
                while (mm) {
                    const int tt = __ffs(mm) - 1;
                    mm &= mm - 1;
                    const float2 g2 = gc[tt];
                    step_row(tb0 + tt * TSTRIDE, g2.x, g2.y, d, e, r);
                }
            }
        }
        }
        if (cur.blk == NBLK - 1) {
            flush_d();
            V *tp = reinterpret_cast<V *>(trace + (size_t)cur.b * AF);
#pragma unroll
            for (int j = 0; j < CH; ++j) {
                if (own[j]) {
#pragma unroll
                    for (int r = 0; r < SCG_A; ++r) tp[r * NCHR + tid + j * NT] = e[j][r];
                }
            }
        }
        __syncthreads();   // the tables just built become readable, the ones just read become writable
        if (nxt.blk == 0) epar ^= 1;
        par ^= 1;
        cur = nxt;
        nxt = nx2;
    }
    float *out = partial + (size_t)blockIdx.x * K * AF;
    for (int i = tid; i < K * AF; i += NTT) out[i] = acc[i];
}

// dW[j] += sum over the per-CTA slabs.  Block (x, y) sums every SCG_RED_SLICES-th slab, four loads in flight per thread
// (the slabs were just written and mostly sit in L2).  Fast mode (default): one atomicAdd per address per slice, so the
// order of the last additions is not fixed and weights agree run-to-run to rounding.  Deterministic mode
// (scg_ctx_set_deterministic; 1.3 -> 2.2 us per step at configs[1]): the slice sums go to a small scratch and the last
// block of column x to finish - an integer ticket per column - adds them in a fixed order: same inputs, same bits.
#define SCG_RED_SLICES 96
__global__ void __launch_bounds__(256) k_reduce(int n_partials, int n, const float *__restrict__ partial, float *red,
                                                unsigned int *tickets, float *__restrict__ dW) {
    __shared__ bool last;
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    const int gy = gridDim.y;
    float mine = 0.f;
    if (j < n) {
        float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;
        int p = blockIdx.y;
        for (; p + 3 * gy < n_partials; p += 4 * gy) {
            s0 += partial[(size_t)p * n + j];
            s1 += partial[(size_t)(p + gy) * n + j];
            s2 += partial[(size_t)(p + 2 * gy) * n + j];
            s3 += partial[(size_t)(p + 3 * gy) * n + j];
        }
        for (; p < n_partials; p += gy) s0 += partial[(size_t)p * n + j];
        mine = (s0 + s1) + (s2 + s3);
    }
    if (!tickets) {                              // fast mode: one atomic per address per slice (order not fixed)
        if (j < n) atomicAdd(dW + j, mine);
        return;
    }
    if (j < n) red[(size_t)blockIdx.y * n + j] = mine;
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) {
        const unsigned int t = atomicAdd(tickets + blockIdx.x, 1u);
        last = (t == (unsigned int)gy - 1u);
        if (last) tickets[blockIdx.x] = 0u;      // ready for the next reduction
    }
    __syncthreads();
    if (last && j < n) {
        __threadfence();
        float t0 = 0.f, t1 = 0.f, t2 = 0.f, t3 = 0.f;     // fixed association: four interleaved chains, then a fixed tree
        int y = 0;
        for (; y + 3 < gy; y += 4) {
            t0 += __ldcg(red + (size_t)y * n + j);
            t1 += __ldcg(red + (size_t)(y + 1) * n + j);
            t2 += __ldcg(red + (size_t)(y + 2) * n + j);
            t3 += __ldcg(red + (size_t)(y + 3) * n + j);
        }
        for (; y < gy; ++y) t0 += __ldcg(red + (size_t)y * n + j);
        dW[j] += (t0 + t1) + (t2 + t3);
    }
}

// slabs -> dW on stream st
static int reduce_slabs(scg_ctx *ctx, int n_slabs, int n, float *dW, cudaStream_t st, bool overlap_prev = false) {
    const int nx_max = (ctx->K * SCG_A * ctx->F + 255) / 256;
    if (ctx->deterministic && !ctx->d_red) {
        SCG_CUDA_OK(cudaMalloc((void **)&ctx->d_red, (size_t)SCG_RED_SLICES * ctx->K * SCG_A * ctx->F * sizeof(float)));
        SCG_CUDA_OK(cudaMalloc((void **)&ctx->d_tickets, (size_t)nx_max * sizeof(unsigned int)));
        SCG_CUDA_OK(cudaMemsetAsync(ctx->d_tickets, 0, (size_t)nx_max * sizeof(unsigned int), st));
    }
    dim3 g((n + 255) / 256, std::min(n_slabs, SCG_RED_SLICES));
    // overlap_prev: the kernel launched just before this one (the example-ring pass) is independent of the reduction, so
    // the reduction may start under it (programmatic dependent launch; k_ring releases its dependents at its start)
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = g;
    cfg.blockDim = dim3(256);
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = overlap_prev ? 1 : 0;
    SCG_CUDA_OK(cudaLaunchKernelEx(&cfg, k_reduce, n_slabs, n, (const float *)ctx->d_partial, ctx->d_red,
                                   ctx->deterministic ? ctx->d_tickets : (unsigned int *)nullptr, dW));
    ++g_scg_launches;
    return 0;
}

__global__ void k_make_rec(int B, const float *__restrict__ x, const float *__restrict__ y,
                           const float *__restrict__ vx, const float *__restrict__ vy, const int *__restrict__ a,
                           const int *__restrict__ option, const float *__restrict__ delta,
                           const uint8_t *__restrict__ done, const uint8_t *__restrict__ mask, int K,
                           float4 *__restrict__ rec, int *__restrict__ cnt) {
    for (int b = blockIdx.x * blockDim.x + threadIdx.x; b < B; b += gridDim.x * blockDim.x) {
        bool act = !mask || mask[b];
        float2 z[4];
        scg_phasors(x[b], y[b], vx[b], vy[b], z);
        int o = min(max(option[b], 0), K - 1);
        uint32_t meta = (uint32_t)(a[b] & 7) | ((uint32_t)o << 8) | (done[b] ? SCG_META_ZERO_AFTER : 0u) |
                        (act ? SCG_META_ACTIVE : 0u);
        rec[(size_t)b * 3 + 0] = make_float4(z[0].x, z[0].y, z[1].x, z[1].y);
        rec[(size_t)b * 3 + 1] = make_float4(z[2].x, z[2].y, z[3].x, z[3].y);
        rec[(size_t)b * 3 + 2] = make_float4(delta[b], __uint_as_float(meta), 0.f, 0.f);
        if (act) atomicAdd(cnt + o, 1);
    }
}

// W += (alpha * alpha_scale_f) * (dW * (steps / cnt_k));  refresh the packed copy;  dW <- 0
// Slots k >= K_opt hold the top-level learner (oracle/option.py): step size alpha_top, mean over the window's events.
template <int N1>
__global__ void k_apply(int K, int K_opt, float *__restrict__ W, float *__restrict__ Wt, float *__restrict__ dW,
                        int *cnt, float alpha, float alpha_top, float steps, unsigned int *ticket) {
    constexpr int F = N1 * N1 * N1 * N1;
    int n = K * SCG_A * F;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        int f = i % F, ka = i / F, a = ka % SCG_A, k = ka / SCG_A;
        int c = cnt[k];
        float w = W[i];
        if (c > 0) {
            int d = f, ss = 0;
#pragma unroll
            for (int j = 0; j < 4; ++j) { int q = d % N1; ss += q * q; d /= N1; }
            float as = (ss == 0) ? 1.0f : (float)(1.0 / sqrt((double)ss));
            const bool top = k >= K_opt;
            float scale = __fdiv_rn(top ? 1.0f : steps, (float)c);
            w = __fadd_rn(w, __fmul_rn(__fmul_rn(top ? alpha_top : alpha, as), __fmul_rn(dW[i], scale)));
            W[i] = w;
        }
        Wt[(size_t)k * WtLayout<N1>::SLOT_FLOATS + WtLayout<N1>::index(a, f)] = w;
        dW[i] = 0.f;
    }
    // the last block to finish zeroes cnt for the next window (every block has read it by then): no extra launch
    __shared__ bool last;
    __syncthreads();
    if (threadIdx.x == 0) {
        __threadfence();
        last = atomicInc(ticket, gridDim.x - 1) == gridDim.x - 1;
    }
    __syncthreads();
    if (last && threadIdx.x < K) cnt[threadIdx.x] = 0;
}

// ---- launch plumbing ----------------------------------------------------------------------------
static int ensure_partials(scg_ctx *ctx, int n);
template <int N1, int VEC, int CPT, int NT, int U>
static int launch_trace_t(scg_ctx *ctx, int B, const float4 *rec, float *trace, float gl, cudaStream_t st) {
    size_t smem = (size_t)ctx->K * SCG_A * ctx->F * sizeof(float);
    if (smem > 227 * 1024) return SCG_ELIMIT;   // the CTA accumulator must fit in shared memory
    auto kern = k_trace<N1, VEC, CPT, NT, U>;
    static ScgKernelCfg cfgc = {};
    int per_sm = 0;
    int rcc = scg_configure(cfgc, kern, NT, smem, &per_sm);
    if (rcc) return rcc;
    if (per_sm < 1) return SCG_ELIMIT;
    int grid = std::max(std::min(B, SCG_NUM_SMS * per_sm), 1);
    int rc = ensure_partials(ctx, grid);
    if (rc) return rc;
    kern<<<grid, NT, smem, st>>>(B, ctx->K, rec, trace, ctx->d_partial, gl);
    SCG_LAUNCH_CHECK();
    return grid;
}

int scg_launch_trace(scg_ctx *ctx, int B, const float *rec, float *trace, float gl, float *dW, cudaStream_t st) {
    int grid = 0;
    const float4 *r4 = reinterpret_cast<const float4 *>(rec);
    static int u3 = -1;
    if (u3 < 0) {   // SCG_TRACE_U: envs in flight per CTA for the order-3 sweep (tuning knob; default 4)
        const char *ev = getenv("SCG_TRACE_U");
        u3 = ev ? atoi(ev) : 4;
    }
    // (A bulk-TMA ring variant of this sweep - cp.async.bulk loads into a shared-memory ring, in-place update, bulk
    // stores - was measured slower on B200: 0.165 vs 0.129 ms at order 3, B = 65,536; with 5 KiB blocks the mbarrier
    // round trips cost more than register-held loads save.  Removed.)
    switch (ctx->order) {
        case 1: grid = launch_trace_t<2, 4, 1, 32, 4>(ctx, B, r4, trace, gl, st); break;
        case 2: grid = launch_trace_t<3, 1, 2, 224, 2>(ctx, B, r4, trace, gl, st); break;
        case 3:
            if (u3 == 2) grid = launch_trace_t<4, 4, 1, 320, 2>(ctx, B, r4, trace, gl, st);
            else if (u3 == 8) grid = launch_trace_t<4, 4, 1, 320, 8>(ctx, B, r4, trace, gl, st);
            else grid = launch_trace_t<4, 4, 1, 320, 4>(ctx, B, r4, trace, gl, st);
            break;
        case 4: grid = launch_trace_t<5, 1, 4, 800, 2>(ctx, B, r4, trace, gl, st); break;
        case 5: grid = launch_trace_t<6, 4, 2, 832, 2>(ctx, B, r4, trace, gl, st); break;
        default: return SCG_ELIMIT;
    }
    if (grid <= 0) return grid == 0 ? SCG_EINVAL : grid;
    return reduce_slabs(ctx, grid, ctx->K * SCG_A * ctx->F, dW, st);
}

static int ensure_partials(scg_ctx *ctx, int n) {
    if (n <= ctx->n_partials && ctx->d_partial) return 0;
    if (ctx->d_partial) cudaFree(ctx->d_partial);
    ctx->d_partial = nullptr;
    ctx->n_partials = 0;
    size_t slab = (size_t)ctx->K * SCG_A * ctx->F * sizeof(float);
    SCG_CUDA_OK(cudaMalloc((void **)&ctx->d_partial, slab * n));
    ctx->n_partials = n;
    return 0;
}

template <int N1>
static size_t window_smem(const scg_ctx *ctx, int k_used, bool multi, bool list = true) {
    constexpr int NN = N1 * N1;
    return (((size_t)k_used * SCG_A * ctx->F + 3) & ~(size_t)3) * sizeof(float) +
           (size_t)2 * SCG_WIN_TB * (N1 == 6 ? NN + 64 : 2 * NN) * sizeof(float2) +
           2 * (multi ? sizeof(WinCtlT<SCG_WIN_MAX>) : (list ? sizeof(WinCtlList) : sizeof(WinCtlT<SCG_WIN_TB>))) + 36 * sizeof(float);
}

// k_used: options that can appear in the window's records (ids 0 .. k_used-1): the CTA accumulator, the slabs and
// the reduction only cover those, which is what lets two CTAs share an SM at order 5 while the chain is short
template <int N1, int VEC, int NT, int CH, bool MULTI, int MINB, bool CTRL, bool LISTP, bool L2PF, bool SPLITP>
static int launch_window_tm(scg_ctx *ctx, int B, int T, int k_used, const float4 *rec, float *trace, float gl,
                            float *dW, cudaStream_t st) {
    const size_t smem = window_smem<N1>(ctx, k_used, MULTI, LISTP);
    auto kern = k_window<N1, VEC, NT, CH, MULTI, MINB, CTRL, LISTP, L2PF, SPLITP>;
    constexpr int NTT = NT + (CTRL ? 32 : 0);
    static ScgKernelCfg cfgc = {};
    int per_sm = 0;
    int rcc = scg_configure(cfgc, kern, NTT, smem, &per_sm);
    if (rcc) return rcc;
    if (per_sm < 1) return SCG_ELIMIT;
    static int cap = -1;
    if (cap < 0) { const char *e = getenv("SCG_WIN_CTAS_PER_SM"); cap = e ? atoi(e) : 0; }
    int occ = cap > 0 ? std::min(cap, per_sm) : per_sm;
    int grid = std::max(1, std::min(B, SCG_NUM_SMS * occ));
    int rc = ensure_partials(ctx, grid);
    if (rc) return rc;
    kern<<<grid, NTT, smem, st>>>(B, k_used, T, rec, trace, ctx->d_partial, gl, dW, ctx->K);
    SCG_LAUNCH_CHECK();
    return grid;
}

template <int N1, int VEC, int NT, int CH, int MINB = 1, bool CTRL = false, bool LISTP = true, bool L2PF = false, bool SPLITP = false>
static int launch_window_t(scg_ctx *ctx, int B, int T, int k_used, const float4 *rec, float *trace, float gl,
                           float *dW, cudaStream_t st) {
    if (T <= SCG_WIN_TB) return launch_window_tm<N1, VEC, NT, CH, false, MINB, CTRL, LISTP, L2PF, SPLITP>(ctx, B, T, k_used, rec, trace, gl, dW, st);
    return launch_window_tm<N1, VEC, NT, CH, true, MINB, CTRL, false, L2PF, false>(ctx, B, T, k_used, rec, trace, gl, dW, st);
}

int scg_prof_push(scg_ctx *ctx, int kind, cudaStream_t st, bool end);

// fold the T recorded steps of the window into the traces and the per-CTA dW slabs; returns the number of slabs (> 0)
// or an error (<= 0); scg_reduce_window then folds the slabs into dW
int scg_launch_window(scg_ctx *ctx, int B, int T, int k_used, const float *rec, float *trace, float gl, float *dW,
                      cudaStream_t st) {
    if (T < 1 || T > SCG_WIN_MAX || k_used < 1 || k_used > ctx->K) return SCG_EINVAL;
    const float4 *r4 = reinterpret_cast<const float4 *>(rec);
    int grid = 0, rc;
    // tuning knobs (defaults = the measured best)
    static int mode3 = -1, list5 = -1;
    if (mode3 < 0) { const char *e = getenv("SCG_WIN_MODE3"); mode3 = e ? atoi(e) : 3; }
    if (list5 < 0) { const char *e = getenv("SCG_WIN_LIST5"); list5 = e ? atoi(e) : 0; }
    if ((rc = scg_prof_push(ctx, 1, st, false))) return rc;
    switch (ctx->order) {
        // <N1, floats per chunk, threads, chunks per thread, min CTAs per SM, scan warp, step lists, L2 prefetch>
        case 1: grid = launch_window_t<2, 4, 32, 1>(ctx, B, T, k_used, r4, trace, gl, dW, st); break;
        case 2: grid = launch_window_t<3, 1, 96, 1>(ctx, B, T, k_used, r4, trace, gl, dW, st); break;
        case 3:
            // two warps per env.  With the next trace prefetched into L2 instead of registers the sweep needs 140
            // registers (7 CTAs per SM) and loses nothing when squeezed to 128 (8 CTAs, while 8 accumulators fit)
            if (mode3 == 3 && 8 * (window_smem<4>(ctx, k_used, T > SCG_WIN_TB) + 1024) <= 227 * 1024)
                grid = launch_window_t<4, 4, 64, 1, 8, false, true, true, true>(ctx, B, T, k_used, r4, trace, gl, dW, st);
            else if (mode3 >= 2 && 8 * (window_smem<4>(ctx, k_used, T > SCG_WIN_TB) + 1024) <= 227 * 1024)
                grid = launch_window_t<4, 4, 64, 1, 8, false, true, true>(ctx, B, T, k_used, r4, trace, gl, dW, st);
            else if (mode3 >= 1) grid = launch_window_t<4, 4, 64, 1, 1, false, true, true>(ctx, B, T, k_used, r4, trace, gl, dW, st);
            else grid = launch_window_t<4, 4, 64, 1>(ctx, B, T, k_used, r4, trace, gl, dW, st);
            break;
        case 4: grid = launch_window_t<5, 1, 640, 1>(ctx, B, T, k_used, r4, trace, gl, dW, st); break;
        case 5:
            // one CTA per SM: 11 work warps + a warp that only runs the per-env scan; next trace prefetched in registers
            // (L2 prefetch: 1.98 vs 1.75 ms - one CTA cannot cover the L2 latency); bit-mask loop (step lists: 1.74 vs 1.70 ms)
            if (list5) grid = launch_window_t<6, 4, 352, 1, 1, true, true>(ctx, B, T, k_used, r4, trace, gl, dW, st);
            else grid = launch_window_t<6, 4, 352, 1, 1, true, false>(ctx, B, T, k_used, r4, trace, gl, dW, st);
            break;
        default: return SCG_ELIMIT;
    }
    if (grid <= 0) return grid == 0 ? SCG_EINVAL : grid;
    if ((rc = scg_prof_push(ctx, 1, st, true))) return rc < 0 ? rc : -rc;
    return grid;
}

int scg_reduce_window(scg_ctx *ctx, int n_slabs, int k_used, float *dW, cudaStream_t st, bool overlap_prev) {
    int rc;
    if ((rc = scg_prof_push(ctx, 2, st, false))) return rc;
    if ((rc = reduce_slabs(ctx, n_slabs, k_used * SCG_A * ctx->F, dW, st, overlap_prev && !ctx->prof_on))) return rc;
    return scg_prof_push(ctx, 2, st, true);
}

extern "C" int scg_ctx_create(int order, int K, scg_ctx_t **out) {
    if (!out) return SCG_EINVAL;
    if (order < 1 || order > SCG_MAX_ORDER || K < 1 || K > SCG_MAX_OPTIONS) return SCG_ELIMIT;
    scg_ctx *c = (scg_ctx *)calloc(1, sizeof(scg_ctx));
    if (!c) return SCG_ENOMEM;
    c->order = order; c->K = K; c->F = scg_pow4(order + 1);
    // (the dense per-step operator scg_sarsa_update keeps a [K][A*F] accumulator per CTA and refuses sizes beyond
    // shared memory when it is called; the agent pipeline's sweep only accumulates the option slots in use)
    *out = c;   // the per-CTA dW slabs are allocated by the first sweep (ensure_partials)
    return 0;
}

extern "C" int scg_ctx_set_deterministic(scg_ctx_t *c, int on) {
    if (!c) return SCG_EINVAL;
    c->deterministic = on ? 1 : 0;
    return 0;
}

extern "C" int scg_ctx_destroy(scg_ctx_t *c) {
    if (!c) return 0;
    if (c->d_partial) cudaFree(c->d_partial);
    if (c->d_rec) cudaFree(c->d_rec);
    if (c->d_red) cudaFree(c->d_red);
    if (c->d_tickets) cudaFree(c->d_tickets);
    if (c->host_ev) cudaEventDestroy(c->host_ev);
    for (int i = 0; i < 2; ++i) if (c->host_st[i]) cudaStreamDestroy(c->host_st[i]);
    for (int i = 0; i < 2 * SCG_HOST_PARTS_MAX + 1; ++i) if (c->host_evs[i]) cudaEventDestroy(c->host_evs[i]);
    if (c->h_ctl) cudaFreeHost(c->h_ctl);
    if (c->d_ring) cudaFree(c->d_ring);
    for (int i = 0; i < 2 * c->prof_cap; ++i) cudaEventDestroy(c->prof_ev[i]);
    free(c->prof_ev);
    free(c->prof_kind);
    free(c);
    return 0;
}

extern "C" int scg_sarsa_update(scg_ctx_t *ctx, int B, const float *x, const float *y, const float *vx,
                                const float *vy, const int *a, const int *option, const float *delta,
                                const uint8_t *done, const uint8_t *mask, float gamma_lambda, float *trace, float *dW,
                                int *cnt, void *stream) {
    if (!ctx || B < 0) return SCG_EINVAL;
    if (B == 0) return 0;
    if (!x || !y || !vx || !vy || !a || !option || !delta || !done || !trace || !dW || !cnt) return SCG_EINVAL;
    cudaStream_t st = (cudaStream_t)stream;
    if (B > ctx->rec_capacity) {
        if (ctx->d_rec) cudaFree(ctx->d_rec);
        ctx->d_rec = nullptr; ctx->rec_capacity = 0;
        SCG_CUDA_OK(cudaMalloc((void **)&ctx->d_rec, (size_t)B * SCG_REC_FLOATS * sizeof(float)));
        ctx->rec_capacity = B;
    }
    int grid = std::max(1, std::min((B + 255) / 256, SCG_NUM_SMS * 8));
    k_make_rec<<<grid, 256, 0, st>>>(B, x, y, vx, vy, a, option, delta, done, mask, ctx->K,
                                     reinterpret_cast<float4 *>(ctx->d_rec), cnt);
    SCG_LAUNCH_CHECK();
    return scg_launch_trace(ctx, B, ctx->d_rec, trace, gamma_lambda, dW, st);
}

extern "C" int scg_apply(int order, int K, float *W, float *Wt, float *dW, int *cnt, float alpha, int window_steps,
                         void *stream) {
    return scg_apply_top(order, K, K, W, Wt, dW, cnt, alpha, 0.f, window_steps, stream);
}

extern "C" int scg_apply_top(int order, int K, int K_opt, float *W, float *Wt, float *dW, int *cnt, float alpha,
                             float alpha_top, int window_steps, void *stream) {
    if (K < 1 || K > SCG_MAX_OPTIONS) return SCG_ELIMIT;
    if (!W || !Wt || !dW || !cnt || K_opt < 0 || K_opt > K) return SCG_EINVAL;
    cudaStream_t st = (cudaStream_t)stream;
    float steps = (float)std::max(window_steps, 1);
    int F = scg_pow4(order + 1);
    int grid = std::max(1, std::min((K * SCG_A * F + 255) / 256, SCG_NUM_SMS * 8));
    // one "blocks done" counter per (device, stream): launches on one stream are ordered, so the counter (which atomicInc
    // wraps back to 0) is reusable launch after launch; applies running concurrently on other streams have their own
    struct Slot { int dev; cudaStream_t st; unsigned int *ticket; };
    static Slot slots[64];
    static int n_slots = 0;
    int dev = 0;
    SCG_CUDA_OK(cudaGetDevice(&dev));
    unsigned int *ticket = nullptr;
    for (int i = 0; i < n_slots; ++i)
        if (slots[i].dev == dev && slots[i].st == st) ticket = slots[i].ticket;
    if (!ticket) {
        if (n_slots >= 64) {           // many short-lived streams: wait for everything, start the table over
            SCG_CUDA_OK(cudaDeviceSynchronize());
            for (int i = 0; i < n_slots; ++i) cudaFree(slots[i].ticket);
            n_slots = 0;
        }
        SCG_CUDA_OK(cudaMalloc((void **)&ticket, sizeof(unsigned int)));
        SCG_CUDA_OK(cudaMemsetAsync(ticket, 0, sizeof(unsigned int), st));
        slots[n_slots++] = Slot{dev, st, ticket};
    }
    DISPATCH_ORDER(order, k_apply<N1><<<grid, 256, 0, st>>>(K, K_opt, W, Wt, dW, cnt, alpha, alpha_top, steps, ticket));
    SCG_LAUNCH_CHECK();
    return 0;
}
