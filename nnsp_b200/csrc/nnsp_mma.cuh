/* nnsp_mma.cuh -- NeuralNetClass_exe for 16 streams per warp on the tensor cores (int8 IMMA).
 *
 * Same reference arithmetic as nnsp_net.cuh (affine.c:12-259, 348-407; lstm.c:15-214; activation.c), but
 * the contraction is formulated across streams: D[stream][unit] = X[stream][k] . W[unit][k], M = 16
 * streams per warp, N = 8 units per tile, K = 32 per instruction (mma.sync.m16n8k32, SASS IMMA.16832).
 * The int16 activations are split exactly into a signed high byte and an unsigned low byte,
 *       x = 256 * hi + lo,   hi = x >> 8 (s8),   lo = x & 0xff (u8),
 * two MMAs (s8 x s8 and u8 x s8) accumulate in int32 and  acc = (acc_hi << 8) + acc_lo  reproduces the
 * reference's accumulator: exactly for the 64-bit-accumulator models (|sum| < 2^31 because cols <= 480),
 * and modulo 2^32 for ACC32BIT_OPT models, which is what __SMLAD computes (affine_acc32b.c:90-101).
 *
 * Why warp-level IMMA and not tcgen05: the GEMMs are tiny (N <= 288, K <= 256) and the kernels are bound by
 * the element-wise finish (LUT activations, LSTM cell), by shared-memory traffic and by latency, not by MMA
 * throughput. tools/tc5_gemm_bench.cu holds both formulations of the layer-0 contraction side by side, bit-exact
 * against a 64-bit reference: tcgen05.mma kind::i8 (128-row tiles, TMEM accumulators, TMA-fed, warp-specialised)
 * runs the S2I shape in 283 us where this mma.sync formulation takes 380 us, with the tensor pipe 10 % busy and the
 * shared-memory pipe 63 % (profiles/r2_tc5_gemm_bench.txt, r2_tc5_ncu.txt); the recurrent scan keeps 16-stream tiles
 * (256 CTAs at 4 096 streams where 128-row tiles would leave 32). IMMA.16832.S8 has a ~100-cycle dependent-issue
 * latency, so k-steps are split over independent accumulator chains.
 *
 * Activations live in per-warp shared memory as two byte planes [16 streams][pitch]; an A fragment
 * register is one aligned 32-bit word of a plane (pitch/4 = 4 mod 8 words -> conflict-free), the epilogue
 * writes byte pairs straight back into the planes of the next layer. Weights are pre-packed on the host
 * into B-fragment order (one 8-byte load per lane per MMA pair). */
#pragma once
#include "nnsp_device.cuh"
#include "nnsp_engine.cuh"
#include "nnsp_net.cuh"

namespace nnsp {

struct MmaLayer {
    int type, act, rows, cols, acc32;
    int kt, ktr;              /* 32-wide k-steps of the input part / of the recurrent part            */
    int nt;                   /* fc: 8-unit tiles; lstm: 8-unit groups (4 gate tiles each)            */
    int sh_x, sh_bias, sh_out;
    int fast;                 /* 1: pre = (acc + bias32) >> -sh_out is exact (no clamp can fire)      */
    int w_off, wh_off;        /* uint2 offsets into the fragment image                                */
    int bias_off;             /* int32 offset into the bias image (canonical row order)               */
};

struct MmaModel {
    int      nn_id, numlayers, act_stride, h_stride, n_out, feat_rshift;
    int      frag_count;      /* uint2 elements in the fragment image (multiple of 2)                 */
    int      bias_count;      /* int32 elements                                                        */
    int      pa;              /* byte pitch of the activation / h planes                               */
    int      hc;              /* h_stride rounded up to 8                                              */
    int      no;              /* n_out rounded up to 8                                                 */
    int      warp_bytes;      /* shared memory per warp                                                */
    int32_t  mean[40], stdR[40];
    int16_t  silence[40];
    MmaLayer layer[NNSP_B200_MAX_LAYERS];
};

/* Context planes are a ring of 12 rows x 40 bytes per stream: the 6-row window starts at row `ctx_base`
 * (byte 40*ctx_base) and slides by one row per frame without moving data; every 7th frame the 5 rows that
 * survive are copied back to the front. Pitch 496 B = 480 + k-step over-read, 124 words = 4 mod 8. */
constexpr int MMA_PC = 496;
constexpr int MMA_RING_ROWS = 12;

__device__ __forceinline__ void imma_s8s8(int (&c)[4], const uint32_t (&a)[4], uint2 b)
{
    asm volatile("mma.sync.aligned.m16n8k32.row.col.s32.s8.s8.s32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+r"(c[0]), "+r"(c[1]), "+r"(c[2]), "+r"(c[3])
                 : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b.x), "r"(b.y));
}
__device__ __forceinline__ void imma_u8s8(int (&c)[4], const uint32_t (&a)[4], uint2 b)
{
    asm volatile("mma.sync.aligned.m16n8k32.row.col.s32.u8.s8.s32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+r"(c[0]), "+r"(c[1]), "+r"(c[2]), "+r"(c[3])
                 : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b.x), "r"(b.y));
}

/* first k-step of an accumulation: D = A.B + C with C given apart from D (zeros, or the bias pair of the lane's two
 * columns repeated for its two rows), so no accumulator has to be initialised by moves */
__device__ __forceinline__ void imma_s8s8_first(int (&d)[4], const uint32_t (&a)[4], uint2 b)
{
    asm volatile("mma.sync.aligned.m16n8k32.row.col.s32.s8.s8.s32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%10,%10,%10,%10};"
                 : "=r"(d[0]), "=r"(d[1]), "=r"(d[2]), "=r"(d[3])
                 : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b.x), "r"(b.y), "r"(0));
}
__device__ __forceinline__ void imma_u8s8_first(int (&d)[4], const uint32_t (&a)[4], uint2 b, int c0, int c1)
{
    asm volatile("mma.sync.aligned.m16n8k32.row.col.s32.u8.s8.s32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%10,%11,%10,%11};"
                 : "=r"(d[0]), "=r"(d[1]), "=r"(d[2]), "=r"(d[3])
                 : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b.x), "r"(b.y), "r"(c0), "r"(c1));
}

/* A fragment of one byte plane: rows g and g+8, k bytes 4q..4q+3 and 16+4q..16+4q+3 of the k-step */
__device__ __forceinline__ void load_a(const uint8_t *plane, int pitch, int kbyte, int g, int q, uint32_t (&a)[4])
{
    const uint8_t *p0 = plane + g * pitch + kbyte + 4 * q;
    const uint8_t *p1 = p0 + 8 * pitch;
    a[0] = *reinterpret_cast<const uint32_t *>(p0);
    a[1] = *reinterpret_cast<const uint32_t *>(p1);
    a[2] = *reinterpret_cast<const uint32_t *>(p0 + 16);
    a[3] = *reinterpret_cast<const uint32_t *>(p1 + 16);
}

/* The same fragment by one ldmatrix.x4. Seen as 8x8 matrices of b16, matrix m covers rows 8(m&1).. and k bytes 16(m>>1)..
 * of the k-step, and thread (g, q) receives bytes 4q..4q+3 of row g of each: exactly a[0..3] above. Lane l supplies the
 * address of row l & 15 at k byte 16 (l >> 4); rows must be 16-byte aligned (every plane pitch is a multiple of 16 with
 * pitch/16 odd, so the eight rows of a matrix fall on distinct bank groups). */
__device__ __forceinline__ uint32_t ldm_lane_addr(const uint8_t *plane, int pitch, int lane)
{
    return (uint32_t)__cvta_generic_to_shared(plane) + (uint32_t)((lane & 15) * pitch + 16 * (lane >> 4));
}
__device__ __forceinline__ void load_a_ldm(uint32_t addr, uint32_t (&a)[4])
{
    asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];"
                 : "=r"(a[0]), "=r"(a[1]), "=r"(a[2]), "=r"(a[3]) : "r"(addr) : "memory");
}

/* two adjacent units of one stream -> byte pair in the high plane and in the low plane */
__device__ __forceinline__ void store_pair(uint8_t *hi, uint8_t *lo, int off, int y0, int y1)
{
    *reinterpret_cast<uint16_t *>(hi + off) = (uint16_t)__byte_perm((uint32_t)y0, (uint32_t)y1, 0x0051);   /* y0.b1 | y1.b1 << 8 */
    *reinterpret_cast<uint16_t *>(lo + off) = (uint16_t)__byte_perm((uint32_t)y0, (uint32_t)y1, 0x0040);   /* y0.b0 | y1.b0 << 8 */
}

/* byte offsets of the tile's buffers from the tile base in shared memory (offsets, not pointers, so that
 * the compiler keeps every access in the shared address space) */
struct WarpPlanes {
    int ctx_hi, ctx_lo;                /* [16][MMA_PC]                                     */
    int act_hi[2], act_lo[2];          /* ping-pong layer outputs, [16][pa]                */
    int h_hi[2], h_lo[2];              /* LSTM h, double buffered (gates read the old h)   */
    int c;                             /* int32 [16][hc]                                   */
    int logits;                        /* int32 [16][no]                                   */
    int scal;                          /* int16 [16][SC_N]                                 */
};

__device__ __forceinline__ WarpPlanes carve_planes(const MmaModel &M)
{
    WarpPlanes w;
    int p = 0;
    w.ctx_hi = p; p += 16 * MMA_PC;
    w.ctx_lo = p; p += 16 * MMA_PC;
    for (int i = 0; i < 2; i++) { w.act_hi[i] = p; p += 16 * M.pa; w.act_lo[i] = p; p += 16 * M.pa; }
    for (int i = 0; i < 2; i++) { w.h_hi[i] = p; p += 16 * M.pa; w.h_lo[i] = p; p += 16 * M.pa; }
    w.c = p; p += 16 * M.hc * 4;
    w.logits = p; p += 16 * M.no * 4;
    w.scal = p;
    return w;
}

/* accumulator pair -> the reference's 32-bit pre-activation (see finish_fc / finish_gate) */
__device__ __forceinline__ int32_t mma_finish(const MmaLayer &L, int32_t acc_x, int32_t acc_h, int32_t bias, bool lstm)
{
    if (L.fast) return (int32_t)((uint32_t)acc_x + (uint32_t)acc_h + (uint32_t)bias) >> (-L.sh_out);
    return lstm ? finish_gate(L, acc_x, acc_h, bias) : finish_fc(L, acc_x, bias);
}

/* One network evaluation for the CTA's 16 streams; the 4 warps split the 8-unit tiles (fc) or the 8-unit
 * groups (lstm) of every layer and meet at a __syncthreads between layers. Input: context planes.
 * `cur` selects the h buffer holding the current state; the new state goes to buffer cur^1 (caller toggles).
 * Weights (B fragments) are read through the read-only L1 path: all CTAs of an SM share one cached copy.
 * tap_act/tap_logits: global bases of the frame row of stream 0 of the tile, stride in elements per stream. */
__device__ __forceinline__ void mma_forward(const MmaModel &M, const uint2 *__restrict__ frag,
                                            const int32_t *__restrict__ bias32, const int16_t *__restrict__ lut,
                                            unsigned char *sm, const WarpPlanes &W, int ctx_k0, int cur,
                                            int warp, int nwarps, int lane, int nvalid,
                                            int16_t *tap_act, long long tap_act_stride,
                                            int32_t *tap_logits, long long tap_logits_stride)
{
    const int g = lane >> 2, q = lane & 3;
    const uint8_t *in_hi = sm + W.ctx_hi, *in_lo = sm + W.ctx_lo;
    int32_t *const cbuf = reinterpret_cast<int32_t *>(sm + W.c);
    int32_t *const logits = reinterpret_cast<int32_t *>(sm + W.logits);
    int in_pitch = MMA_PC, in_k0 = ctx_k0, pp = 0, ho = 0, ao = 0;
    for (int li = 0; li < M.numlayers; li++) {
        const MmaLayer &L = M.layer[li];
        const bool last = (li == M.numlayers - 1);
        const int32_t *B = bias32 + L.bias_off;
        if (L.type == LAYER_LSTM) {
            const int H = L.rows;
            uint8_t *oh = sm + (cur ? W.h_hi[0] : W.h_hi[1]), *ol = sm + (cur ? W.h_lo[0] : W.h_lo[1]);
            const uint8_t *hh = sm + (cur ? W.h_hi[1] : W.h_hi[0]), *hl = sm + (cur ? W.h_lo[1] : W.h_lo[0]);
            for (int grp = warp; grp < L.nt; grp += nwarps) {
                int ax[4][4], ah[4][4];
                {   /* input half: 4 gate tiles share the A fragments (rc_Krows_8x16 first affine) */
                    int ch[4][4] = {}, cl[4][4] = {};
                    const uint2 *wf = frag + L.w_off + (long long)grp * 4 * L.kt * 32 + lane;
                    for (int ks = 0; ks < L.kt; ks++) {
                        uint32_t fh[4], fl[4];
                        load_a(in_hi, in_pitch, in_k0 + 32 * ks, g, q, fh);
                        load_a(in_lo, in_pitch, in_k0 + 32 * ks, g, q, fl);
#pragma unroll
                        for (int gt = 0; gt < 4; gt++) {
                            const uint2 b = __ldg(wf + (gt * L.kt + ks) * 32);
                            imma_s8s8(ch[gt], fh, b);
                            imma_u8s8(cl[gt], fl, b);
                        }
                    }
#pragma unroll
                    for (int gt = 0; gt < 4; gt++)
#pragma unroll
                        for (int e = 0; e < 4; e++) ax[gt][e] = (int)(((uint32_t)ch[gt][e] << 8) + (uint32_t)cl[gt][e]);
                }
                {   /* recurrent half on the OLD h (lstm.c:54-104 all read h_state before :205-206 updates it) */
                    int ch[4][4] = {}, cl[4][4] = {};
                    const uint2 *wf = frag + L.wh_off + (long long)grp * 4 * L.ktr * 32 + lane;
                    for (int ks = 0; ks < L.ktr; ks++) {
                        uint32_t fh[4], fl[4];
                        load_a(hh, M.pa, ho + 32 * ks, g, q, fh);
                        load_a(hl, M.pa, ho + 32 * ks, g, q, fl);
#pragma unroll
                        for (int gt = 0; gt < 4; gt++) {
                            const uint2 b = __ldg(wf + (gt * L.ktr + ks) * 32);
                            imma_s8s8(ch[gt], fh, b);
                            imma_u8s8(cl[gt], fl, b);
                        }
                    }
#pragma unroll
                    for (int gt = 0; gt < 4; gt++)
#pragma unroll
                        for (int e = 0; e < 4; e++) ah[gt][e] = (int)(((uint32_t)ch[gt][e] << 8) + (uint32_t)cl[gt][e]);
                }
                /* gates + cell for the lane's 2 rows x 2 units of this group */
                const int u0 = 8 * grp + 2 * q;
#pragma unroll
                for (int rr = 0; rr < 2; rr++) {
                    const int row = g + 8 * rr;
                    int y[2];
#pragma unroll
                    for (int cc = 0; cc < 2; cc++) {
                        const int e = 2 * rr + cc, u = u0 + cc;
                        int o = 0;
                        if (u < H) {
                            const int32_t gi = sigmoid_q15(mma_finish(L, ax[0][e], ah[0][e], B[u], true), lut);
                            const int32_t gj = tanh_q15(mma_finish(L, ax[1][e], ah[1][e], B[H + u], true), lut);
                            const int32_t gf = sigmoid_q15(mma_finish(L, ax[2][e], ah[2][e], B[2 * H + u], true), lut);
                            const int32_t go = sigmoid_q15(mma_finish(L, ax[3][e], ah[3][e], B[3 * H + u], true), lut);
                            int32_t *cp = cbuf + row * M.hc + ho + u;
                            const int64_t t = ((int64_t)gi * (int64_t)gj + (int64_t)gf * (int64_t)(*cp)) >> 15;   /* lstm.c:108-109 */
                            const int32_t cn = sat32_dev(t);
                            *cp = cn;
                            o = (tanh_q15(cn, lut) * go) >> 15;                                                   /* lstm.c:111-115 */
                            o = o > 32767 ? 32767 : (o < -32768 ? -32768 : o);
                            if (tap_act && !last && row < nvalid) tap_act[row * tap_act_stride + ao + u] = (int16_t)o;
                            if (last) logits[row * M.no + u] = o;
                        }
                        y[cc] = o;
                    }
                    store_pair(oh, ol, row * M.pa + ho + u0, y[0], y[1]);
                }
            }
            in_hi = oh; in_lo = ol; in_pitch = M.pa; in_k0 = ho;
            ho += H;
        } else {
            uint8_t *oh = sm + (pp ? W.act_hi[1] : W.act_hi[0]), *ol = sm + (pp ? W.act_lo[1] : W.act_lo[0]);
            for (int nt = warp; nt < L.nt; nt += nwarps) {
                /* even and odd k-steps accumulate in separate chains (98-cycle IMMA dependent latency) */
                int ch[2][4] = {}, cl[2][4] = {};
                const uint2 *wf = frag + L.w_off + (long long)nt * L.kt * 32 + lane;
                for (int ks = 0; ks < L.kt; ks += 2) {
                    uint32_t fh[4], fl[4];
                    load_a(in_hi, in_pitch, in_k0 + 32 * ks, g, q, fh);
                    load_a(in_lo, in_pitch, in_k0 + 32 * ks, g, q, fl);
                    const uint2 b0 = __ldg(wf + ks * 32);
                    imma_s8s8(ch[0], fh, b0);
                    imma_u8s8(cl[0], fl, b0);
                    if (ks + 1 < L.kt) {
                        load_a(in_hi, in_pitch, in_k0 + 32 * ks + 32, g, q, fh);
                        load_a(in_lo, in_pitch, in_k0 + 32 * ks + 32, g, q, fl);
                        const uint2 b1 = __ldg(wf + (ks + 1) * 32);
                        imma_s8s8(ch[1], fh, b1);
                        imma_u8s8(cl[1], fl, b1);
                    }
                }
                const int nb = nt * 8 + 2 * q;
#pragma unroll
                for (int rr = 0; rr < 2; rr++) {
                    const int row = g + 8 * rr;
                    int y[2];
#pragma unroll
                    for (int cc = 0; cc < 2; cc++) {
                        const int e = 2 * rr + cc, n = nb + cc;
                        int o = 0;
                        if (n < L.rows) {
                            const int32_t acc = (int32_t)((((uint32_t)ch[0][e] + (uint32_t)ch[1][e]) << 8) + (uint32_t)cl[0][e] + (uint32_t)cl[1][e]);
                            const int32_t pre = mma_finish(L, acc, 0, B[n], false);
                            if (L.act == ACT_LINEAR) { logits[row * M.no + n] = pre; o = 0; }      /* activation.c:19-29 */
                            else {
                                o = activate16(L.act, pre, lut);
                                if (last) logits[row * M.no + n] = o;                                 /* neural_nets.c:160-166 */
                                else if (tap_act && row < nvalid) tap_act[row * tap_act_stride + ao + n] = (int16_t)o;
                            }
                        }
                        y[cc] = o;
                    }
                    store_pair(oh, ol, row * M.pa + nb, y[0], y[1]);
                }
            }
            in_hi = oh; in_lo = ol; in_pitch = M.pa; in_k0 = 0;
            pp ^= 1;
        }
        if (!last) ao += L.rows;
        __syncthreads();
    }
    if (tap_logits)
        for (int i = warp * 32 + lane; i < 16 * M.n_out; i += nwarps * 32) {
            const int row = i / M.n_out, n = i - row * M.n_out;
            if (row < nvalid) tap_logits[row * tap_logits_stride + n] = logits[row * M.no + n];
        }
}

}  // namespace nnsp
