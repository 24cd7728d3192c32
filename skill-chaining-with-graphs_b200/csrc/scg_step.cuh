// scg_step.cuh - device side of the Pinball step, shared by the standalone K1 kernel (scg_step.cu)
// and the fused agent-step kernel (scg_agent.cu).
//
// Semantics follow oracle/pinball.py step_scalar operation for operation: every fp32 +,-,* is an
// explicit round-to-nearest intrinsic so that nvcc cannot contract a*b+c into an FMA and the collision
// decisions, terminal flags and next states are bit-identical to the oracle.
#pragma once
#include "scg_common.cuh"

// ---------------------------------------------------------------------------------------------
// device: TMA bulk staging of the map blob
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void stage_blob(unsigned char *smem, const unsigned char *gblob, int bytes,
                                           unsigned long long *bar) {
    uint32_t bar_a = smem_u32(bar);
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar_a));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar_a), "r"(bytes) : "memory");
        asm volatile(
            "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                smem_u32(smem)),
            "l"(gblob), "r"(bytes), "r"(bar_a)
            : "memory");
    }
    uint32_t ok = 0;
    while (!ok) {
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(ok)
            : "r"(bar_a)
            : "memory");
    }
}

// ---------------------------------------------------------------------------------------------
// device: the step, shared by scg_step and the fused agent step
// ---------------------------------------------------------------------------------------------
struct StepMap {
    const float4 *ea, *eb;
    const uint32_t *cells;
    const uint16_t *cand;
    int n_edges, G;
    float gf, h, r2, tx, ty, tr2;
};

__device__ __forceinline__ StepMap make_step_map(const unsigned char *blob) {
    const ScgMapHeader *h = reinterpret_cast<const ScgMapHeader *>(blob);
    StepMap m;
    m.ea = reinterpret_cast<const float4 *>(blob + h->off_edges_a);
    m.eb = reinterpret_cast<const float4 *>(blob + h->off_edges_b);
    m.cells = reinterpret_cast<const uint32_t *>(blob + h->off_cells);
    m.cand = reinterpret_cast<const uint16_t *>(blob + h->off_cand);
    m.n_edges = h->n_edges; m.G = h->grid_n; m.gf = h->grid_f;
    m.h = h->h; m.r2 = h->r2; m.tx = h->tx; m.ty = h->ty; m.tr2 = h->tr2;
    return m;
}

__device__ __forceinline__ bool edge_hit(float4 ea, float inv, float x, float y, float vx, float vy, float r2) {
    float rx = __fsub_rn(x, ea.x), ry = __fsub_rn(y, ea.y);
    float t = __fmul_rn(__fadd_rn(__fmul_rn(rx, ea.z), __fmul_rn(ry, ea.w)), inv);
    t = fminf(fmaxf(t, 0.f), 1.f);
    float cx = __fadd_rn(ea.x, __fmul_rn(t, ea.z)), cy = __fadd_rn(ea.y, __fmul_rn(t, ea.w));
    float ex = __fsub_rn(cx, x), ey = __fsub_rn(cy, y);
    float d2 = __fadd_rn(__fmul_rn(ex, ex), __fmul_rn(ey, ey));
    float dot = __fadd_rn(__fmul_rn(ex, vx), __fmul_rn(ey, vy));
    return (d2 <= r2) && (dot > 0.f);
}

template <bool kCull>
__device__ __forceinline__ void pinball_step(const StepMap &m, float &x, float &y, float &vx, float &vy, int a,
                                             float &reward, int &flags) {
    const float kImpulse = 0.2f;  // fp32(1)/fp32(5)
    if (a == 0) vx = fminf(fmaxf(__fadd_rn(vx, kImpulse), -1.f), 1.f);
    else if (a == 1) vy = fminf(fmaxf(__fadd_rn(vy, kImpulse), -1.f), 1.f);
    else if (a == 2) vx = fminf(fmaxf(__fsub_rn(vx, kImpulse), -1.f), 1.f);
    else if (a == 3) vy = fminf(fmaxf(__fsub_rn(vy, kImpulse), -1.f), 1.f);
    int kind = 0, ids = 0;
    bool done = false;
#pragma unroll 1
    for (int i = 0; i < 20; ++i) {
        x = __fadd_rn(x, __fmul_rn(vx, m.h));
        y = __fadd_rn(y, __fmul_rn(vy, m.h));
        int nhit = 0, first = -1;
        // cell of (x, y): floor(x G) is exact (G is a power of two), and 0 <= x < 1 is 0 <= floor(x G) < G
        const int ci = __float2int_rd(x * m.gf), cj = __float2int_rd(y * m.gf);
        const bool in_grid = kCull && (unsigned)ci < (unsigned)m.G && (unsigned)cj < (unsigned)m.G;
        bool goal_near = true;
        if (in_grid) {
            const uint32_t cell = m.cells[ci * m.G + cj];
            goal_near = (cell >> 31) != 0u;      // set by scg_map_create where the goal test can succeed
            const int cnt = cell & 0xFF, start = (cell >> 8) & 0x7FFFFF;
#pragma unroll 1   // lists are short (0-4 edges): the unrolled-by-4 prologue cost more than it saved
            for (int j = 0; j < cnt; ++j) {
                int e = m.cand[start + j];
                float4 ea = m.ea[e];
                float inv = m.eb[e].x;
                if (edge_hit(ea, inv, x, y, vx, vy, m.r2)) {
                    if (first < 0) first = e;
                    ++nhit;
                }
            }
        } else {
            for (int e = 0; e < m.n_edges; ++e) {
                float4 ea = m.ea[e];
                float inv = m.eb[e].x;
                if (edge_hit(ea, inv, x, y, vx, vy, m.r2)) {
                    if (first < 0) first = e;
                    ++nhit;
                }
            }
        }
        if (nhit == 1) {
            float4 eb = m.eb[first];
            float k = __fmul_rn(2.0f, __fadd_rn(__fmul_rn(vx, eb.y), __fmul_rn(vy, eb.z)));
            vx = __fsub_rn(vx, __fmul_rn(k, eb.y));
            vy = __fsub_rn(vy, __fmul_rn(k, eb.z));
            kind = 1;
            ids = __float_as_int(eb.w);
            if (i == 19) {
                x = __fadd_rn(x, __fmul_rn(vx, m.h));
                y = __fadd_rn(y, __fmul_rn(vy, m.h));
            }
        } else if (nhit >= 2) {
            vx = -vx;
            vy = -vy;
            kind = 2;
            ids = __float_as_int(m.eb[first].w);
        }
        if (goal_near) {
            float gx = __fsub_rn(x, m.tx), gy = __fsub_rn(y, m.ty);
            if (__fadd_rn(__fmul_rn(gx, gx), __fmul_rn(gy, gy)) < m.tr2) {
                done = true;
                break;
            }
        }
    }
    if (done) {
        reward = 10000.0f;
    } else {
        vx = __fmul_rn(vx, 0.995f);
        vy = __fmul_rn(vy, 0.995f);
        if (x > 1.f) x = 0.95f;
        if (x < 0.f) x = 0.05f;
        if (y > 1.f) y = 0.95f;
        if (y < 0.f) y = 0.05f;
        reward = (a == 4) ? -1.0f : -5.0f;
    }
    flags = (done ? 1 : 0) | (kind << SCG_FLAG_KIND_SHIFT) | (kind ? (ids << SCG_FLAG_EDGE_SHIFT) : 0);
}

