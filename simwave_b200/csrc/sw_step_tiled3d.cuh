// Tiled 3D step kernel for sm_100a (float32, constant density).
//
// Decomposition: a CTA owns a (BX x BY) tile of the (M,F) = (x,y) plane and
// marches along S (= z) over a chunk of planes.
//
//   * u_cur planes, with their halo, are brought into a shared-memory ring by
//     TMA (cp.async.bulk.tensor.3d, one elected thread, completion on an
//     mbarrier).  Out-of-range parts of a box are zero-filled by the TMA unit,
//     so edge tiles need no special loads.  The ring holds the R+1 planes
//     between the centre plane and the newest plane plus PF planes in flight.
//   * The S-direction neighbours live in a per-thread register queue of 2R+1
//     values per point (the classic 2.5D scheme); each thread feeds the queue
//     from the newest plane in the ring.
//   * M- and F-direction neighbours are read from the centre plane in the
//     ring with 128-bit loads; each thread updates a PM x 4 register tile.
//   * u_prev, c0, q stream straight from HBM with 128-bit loads issued one
//     plane ahead; u_next leaves with 128-bit stores.  The damping factors
//     and the boundary conditions are applied in the same kernel
//     (store_with_boundaries semantics).
//
// The arithmetic goes through the same helpers as the plain kernel, in the
// same order, so in strict mode the two kernels (and the reference) agree bit
// for bit.
#pragma once

#include <cuda.h>

#include "sw_math.cuh"
#include "sw_step_simple.cuh"

namespace sw {

// ---- PTX helpers -------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void *p)
{
    return (uint32_t)__cvta_generic_to_shared(p);
}
__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count)
                 : "memory");
}
__device__ __forceinline__ void mbar_fence_init()
{
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)),
                 "r"(bytes)
                 : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity)
{
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "WAIT_%=:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra DONE_%=;\n\t"
        "bra WAIT_%=;\n\t"
        "DONE_%=:\n\t"
        "}" ::"r"(smem_u32(bar)),
        "r"(parity)
        : "memory");
}
__device__ __forceinline__ void fence_proxy_async()
{
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
__device__ __forceinline__ void tma_load_3d(void *dst, const CUtensorMap *map, uint64_t *bar,
                                            int c0, int c1, int c2)
{
    asm volatile(
        "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes"
        " [%0], [%1, {%3, %4, %5}], [%2];" ::"r"(smem_u32(dst)),
        "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2)
        : "memory");
}
// read-only model streams: non-coherent path, do not pollute L1
__device__ __forceinline__ float4 ldg_stream(const float *p)
{
    float4 v;
    asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];"
                 : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w)
                 : "l"(p));
    return v;
}
// u_prev may alias u_next (the reference updates in place between snapshots,
// 3d/wave.c:613-617), so it takes the coherent path
__device__ __forceinline__ float4 ldg_field(const float *p)
{
    float4 v;
    asm volatile("ld.global.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];"
                 : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w)
                 : "l"(p)
                 : "memory");
    return v;
}

// ---- tile geometry -------------------------------------------------------------
template <int R, int PM, int TX, int TY, int PF>
struct Tile3D {
    static constexpr int RP = (R + 3) / 4 * 4;       // F halo rounded to a float4
    static constexpr int BX = TY * PM;               // rows (M) per tile
    static constexpr int BY = TX * 4;                // columns (F) per tile
    static constexpr int BXH = BX + 2 * R;
    static constexpr int BYH = BY + 2 * RP;
    static constexpr int NS = R + 1 + PF;            // ring slots
    static constexpr int SLOT_BYTES = (BXH * BYH * 4 + 127) / 128 * 128;
    static constexpr int SLOT_FLOATS = SLOT_BYTES / 4;
    static constexpr int BOX_BYTES = BXH * BYH * 4;
    static constexpr int SMEM_BYTES = NS * SLOT_BYTES + NS * 8;
    static constexpr int THREADS = TX * TY;
};

// Boundary-aware store of four consecutive F points (f .. f+3) of row (s,m).
// Same semantics as store_with_boundaries, vectorised where nothing special
// happens.  `nvalid` = how many of the four columns are interior points.
__device__ __forceinline__ void store_row4(const StepArgs<float> &a, int s, int m, int f,
                                           const float v[4], int nvalid)
{
    const Grid &g = a.g;
    const int r = g.r;
    const int firstF = r, lastF = g.nF - r - 1;
    const int firstM = r, lastM = g.nM - r - 1;
    const int firstS = r, lastS = g.nS - r - 1;
    float *next = a.next;
    const long long p = g.at(s, m, f);

    const bool zMb = (a.bc[2] == 1) & (m == firstM);
    const bool zMa = (a.bc[3] == 1) & (m == lastM);
    const bool zSb = (a.bc[0] == 1) & (s == firstS);
    const bool zSa = (a.bc[1] == 1) & (s == lastS);
    const bool zRow = zMb | zMa | zSb | zSa;
    bool zFb[4], zFa[4];
    float own[4];
#pragma unroll
    for (int c = 0; c < 4; c++) {
        zFb[c] = (a.bc[4] == 1) & (f + c == firstF);
        zFa[c] = (a.bc[5] == 1) & (f + c == lastF);
        own[c] = (zRow | zFb[c] | zFa[c]) ? 0.0f : v[c];
    }
    if (nvalid == 4) {
        *reinterpret_cast<float4 *>(next + p) = make_float4(own[0], own[1], own[2], own[3]);
    } else {
#pragma unroll
        for (int c = 0; c < 4; c++)
            if (c < nvalid)
                next[p + c] = own[c];
    }

    const bool anyNeumann = (a.bc[0] == 2) | (a.bc[1] == 2) | (a.bc[2] == 2) | (a.bc[3] == 2) |
                            (a.bc[4] == 2) | (a.bc[5] == 2);
    if (!anyNeumann)
        return;

    // F pass
    if (a.bc[4] == 2 && f <= firstF + r) {
#pragma unroll
        for (int c = 0; c < 4; c++) {
            const int fc = f + c;
            if (c < nvalid && fc > firstF && fc <= firstF + r)
                next[p + c - 2 * (fc - firstF)] = v[c];
        }
    }
    if (a.bc[5] == 2 && f + 3 >= lastF - r) {
#pragma unroll
        for (int c = 0; c < 4; c++) {
            const int fc = f + c;
            if (c < nvalid && fc < lastF && fc >= lastF - r)
                next[p + c + 2 * (lastF - fc)] = zFb[c] ? 0.0f : v[c];
        }
    }
    // M pass sees the F pass' zeroing
    const bool mb = a.bc[2] == 2 && m > firstM && m <= firstM + r;
    const bool ma = a.bc[3] == 2 && m < lastM && m >= lastM - r;
    const bool sb = a.bc[0] == 2 && s > firstS && s <= firstS + r;
    const bool sa = a.bc[1] == 2 && s < lastS && s >= lastS - r;
    if (mb | ma | sb | sa) {
#pragma unroll
        for (int c = 0; c < 4; c++) {
            if (c >= nvalid)
                continue;
            const bool zF = zFb[c] | zFa[c];
            const float vM = zF ? 0.0f : v[c];
            if (mb)
                next[p + c - 2 * (long long)(m - firstM) * g.pitch] = vM;
            if (ma)
                next[p + c + 2 * (long long)(lastM - m) * g.pitch] = zMb ? 0.0f : vM;
            const float vS = (zF | zMb | zMa) ? 0.0f : v[c];
            if (sb)
                next[p + c - 2 * (long long)(s - firstS) * g.planeStride] = vS;
            if (sa)
                next[p + c + 2 * (long long)(lastS - s) * g.planeStride] = zSb ? 0.0f : vS;
        }
    }
}

template <int R, int PM, int TX, int TY, int PF, int MATH, int MINB>
__global__ void __launch_bounds__(TX *TY, MINB)
step3d_tiled_kernel(const __grid_constant__ StepArgs<float> a,
                    const __grid_constant__ CUtensorMap mapCur, int zChunk)
{
    using TL = Tile3D<R, PM, TX, TY, PF>;
    constexpr int RP = TL::RP, BYH = TL::BYH, NS = TL::NS;
    constexpr int Q = 2 * R + 1;
    const Grid &g = a.g;

    // TMA destinations must be 128-byte aligned; there is no static shared
    // memory in this kernel, so the dynamic segment starts the window
    extern __shared__ __align__(128) unsigned char smem[];
    float *ring = reinterpret_cast<float *>(smem);
    uint64_t *full = reinterpret_cast<uint64_t *>(smem + NS * TL::SLOT_BYTES);

    const int tid = threadIdx.x;
    const int tx = tid % TX, ty = tid / TX;
    const int f0 = R + blockIdx.x * TL::BY;
    const int m0 = R + blockIdx.y * TL::BX;
    const int z0 = R + blockIdx.z * zChunk;
    const int z1 = min(z0 + zChunk, g.nS - R);
    const int planes = z1 - z0;
    const int L = planes + 2 * R;       // planes streamed: z0-R .. z1+R-1

    if (tid == 0) {
        if (smem_u32(ring) & 127u)
            __trap();
        for (int s = 0; s < NS; s++)
            mbar_init(&full[s], 1);
        mbar_fence_init();
    }
    __syncthreads();

    auto issue = [&](int l) {
        const int slot = l % NS;
        mbar_expect_tx(&full[slot], TL::BOX_BYTES);
        tma_load_3d(ring + slot * TL::SLOT_FLOATS, &mapCur, &full[slot], g.lpad + f0 - RP,
                    m0 - R, z0 - R + l);
    };
    if (tid == 0) {
        const int first = min(NS, L);
        for (int l = 0; l < first; l++)
            issue(l);
    }

    // my points: rows m0 + ty*PM + i, columns f0 + 4*tx .. +3
    const int fMine = f0 + 4 * tx;
    const int lastF = g.nF - R - 1, lastM = g.nM - R - 1;
    int nvalid = lastF - fMine + 1;
    nvalid = nvalid < 0 ? 0 : (nvalid > 4 ? 4 : nvalid);
    bool rowValid[PM];
#pragma unroll
    for (int i = 0; i < PM; i++)
        rowValid[i] = (m0 + ty * PM + i <= lastM) && nvalid > 0;

    const int srow = ty * PM + R;           // my first row inside a ring slot
    const int scol = 4 * tx + RP;           // my first column inside a ring slot

    auto slot_ptr = [&](int l) { return ring + (l % NS) * TL::SLOT_FLOATS; };
    auto lds4 = [&](const float *slot, int row, int col) {
        return *reinterpret_cast<const float4 *>(slot + row * BYH + col);
    };

    // register queue over S: qv[i][c][k] holds plane (centre - R + k)
    float qv[PM][4][Q];

    // prime the queue with planes z0-R .. z0+R-1.  The first R of them are
    // never centre planes, so their slots are refilled as soon as every
    // thread has copied its values out.
#pragma unroll
    for (int l = 0; l < 2 * R; l++) {
        mbar_wait(&full[l % NS], (l / NS) & 1);
        const float *slot = slot_ptr(l);
#pragma unroll
        for (int i = 0; i < PM; i++) {
            const float4 v = lds4(slot, srow + i, scol);
            // positions 1..2R: the loop's shift brings them to 0..2R-1
            qv[i][0][l + 1] = v.x; qv[i][1][l + 1] = v.y;
            qv[i][2][l + 1] = v.z; qv[i][3][l + 1] = v.w;
        }
        if (l < R) {
            __syncthreads();
            if (tid == 0 && l + NS < L) {
                fence_proxy_async();
                issue(l + NS);
            }
        }
    }

    // streams for the first plane
    float4 pv[PM], c0v[PM], qd[PM];
    auto load_streams = [&](int s) {
#pragma unroll
        for (int i = 0; i < PM; i++) {
            if (rowValid[i]) {
                const long long p = g.at(s, m0 + ty * PM + i, fMine);
                pv[i] = ldg_field(a.prev + p);
                c0v[i] = ldg_stream(a.c0 + p);
                qd[i] = ldg_stream(a.q + p);
            }
        }
    };
    load_streams(z0);

    for (int j = 0; j < planes; j++) {
        const int s = z0 + j;
        const int lf = j + 2 * R;       // newest plane needed
        const int lc = j + R;           // centre plane

        // shift the queue and take the newest plane from the ring
        mbar_wait(&full[lf % NS], (lf / NS) & 1);
        {
            const float *slot = slot_ptr(lf);
#pragma unroll
            for (int i = 0; i < PM; i++) {
                const float4 v = lds4(slot, srow + i, scol);
#pragma unroll
                for (int c = 0; c < 4; c++)
#pragma unroll
                    for (int k = 0; k < Q - 1; k++)
                        qv[i][c][k] = qv[i][c][k + 1];
                qv[i][0][Q - 1] = v.x; qv[i][1][Q - 1] = v.y;
                qv[i][2][Q - 1] = v.z; qv[i][3][Q - 1] = v.w;
            }
        }

        const float *ctr = slot_ptr(lc);
        float out[PM][4];

#pragma unroll
        for (int i = 0; i < PM; i++) {
            // F direction: window of 4 + 2*RP values around my four points
            float w[4 + 2 * RP];
#pragma unroll
            for (int b = 0; b < (4 + 2 * RP) / 4; b++) {
                const float4 v = lds4(ctr, srow + i, scol - RP + 4 * b);
                w[4 * b + 0] = v.x; w[4 * b + 1] = v.y; w[4 * b + 2] = v.z; w[4 * b + 3] = v.w;
            }
            float sdF[4], sdM[4], sdS[4];
#pragma unroll
            for (int c = 0; c < 4; c++) {
                const float centre = Ops<float>::mul(a.c2[0], qv[i][c][R]);
                sdF[c] = centre; sdM[c] = centre; sdS[c] = centre;
            }
#pragma unroll
            for (int ir = 1; ir <= R; ir++) {
                const float4 up = lds4(ctr, srow + i + ir, scol);
                const float4 dn = lds4(ctr, srow + i - ir, scol);
                const float upv[4] = {up.x, up.y, up.z, up.w};
                const float dnv[4] = {dn.x, dn.y, dn.z, dn.w};
#pragma unroll
                for (int c = 0; c < 4; c++) {
                    sdF[c] = ring_sum<float, MATH>(sdF[c], a.c2[ir], w[RP + c + ir], w[RP + c - ir]);
                    sdM[c] = ring_sum<float, MATH>(sdM[c], a.c2[ir], upv[c], dnv[c]);
                    sdS[c] = ring_sum<float, MATH>(sdS[c], a.c2[ir], qv[i][c][R + ir],
                                                   qv[i][c][R - ir]);
                }
            }
            const float pvv[4] = {pv[i].x, pv[i].y, pv[i].z, pv[i].w};
            const float c0a[4] = {c0v[i].x, c0v[i].y, c0v[i].z, c0v[i].w};
            const float qa[4] = {qd[i].x, qd[i].y, qd[i].z, qd[i].w};
#pragma unroll
            for (int c = 0; c < 4; c++) {
                const float lap = laplacian<float, 3, MATH>(sdS[c], sdM[c], sdF[c], a.h2, a.inv_h2);
                out[i][c] = leapfrog<float>(lap, qv[i][c][R], pvv[c], c0a[c], qa[c]);
            }
        }

        // everyone is done with the centre plane: refill its slot
        __syncthreads();
        if (tid == 0 && lc + NS < L) {
            fence_proxy_async();
            issue(lc + NS);
        }

        // streams of the next plane go out before this plane's stores
        if (j + 1 < planes)
            load_streams(s + 1);

#pragma unroll
        for (int i = 0; i < PM; i++) {
            if (rowValid[i]) {
                if (a.fuse_bc) {
                    store_row4(a, s, m0 + ty * PM + i, fMine, out[i], nvalid);
                } else {
                    const long long p = g.at(s, m0 + ty * PM + i, fMine);
                    if (nvalid == 4) {
                        *reinterpret_cast<float4 *>(a.next + p) =
                            make_float4(out[i][0], out[i][1], out[i][2], out[i][3]);
                    } else {
                        for (int c = 0; c < nvalid; c++)
                            a.next[p + c] = out[i][c];
                    }
                }
            }
        }
    }
}

}  // namespace sw
