/* nnsp_mma.cu -- host packing + kernel of the tensor-core network path (see nnsp_mma.cuh).
 * nn_mma_kernel: one CTA (4 warps) = one tile of 16 streams, all T frames of the call in order. */
#include <cuda_runtime.h>
#include <mutex>
#include <stdlib.h>
#include <string.h>

#include "nnsp_host.h"
#include "nnsp_mma.cuh"

namespace nnsp {

constexpr int MMA_WARPS = 4;
constexpr int MMA_THREADS = MMA_WARPS * 32;

struct MmaArgs {
    const MmaModel *model;
    const uint2 *frag;
    const int32_t *bias32;
    const DevTables *tables;
    StreamState st;
    const int32_t *logmel;
    int s0, ns, T;
    nnsp_b200_result *results;
    nnsp_b200_taps taps;
    int16_t thresh_prob, th_count;
    int raw_ctx;
};

__global__ void __launch_bounds__(MMA_THREADS, 4)
nn_mma_kernel(MmaArgs a, int off_lut, int off_model, int off_tile, int smem_total)
{
    extern __shared__ __align__(128) unsigned char smem_raw[];
    int32_t *bias32 = reinterpret_cast<int32_t *>(smem_raw);
    int16_t *lut = reinterpret_cast<int16_t *>(smem_raw + off_lut);
    MmaModel &M = *reinterpret_cast<MmaModel *>(smem_raw + off_model);
    const int tid = threadIdx.x;
    {
        const int *src = reinterpret_cast<const int *>(a.model);
        int *dst = reinterpret_cast<int *>(&M);
        for (int i = tid; i < (int)(sizeof(MmaModel) / 4); i += MMA_THREADS) dst[i] = src[i];
        for (int i = tid; i < 384; i += MMA_THREADS) lut[i] = a.tables->tanh_lut[i];
        /* planes start zeroed: padded k columns and padded units must hold finite bytes (they meet zero weights) */
        int *z = reinterpret_cast<int *>(smem_raw + off_tile);
        for (int i = tid; i < (smem_total - off_tile) / 4; i += MMA_THREADS) z[i] = 0;
    }
    __syncthreads();
    for (int i = tid; i < M.bias_count; i += MMA_THREADS) bias32[i] = a.bias32[i];
    __syncthreads();

    const int warp = tid >> 5, lane = tid & 31;
    const WarpPlanes W = carve_planes(M);
    unsigned char *const sm = smem_raw + off_tile;
    uint8_t *const ctx_hi = sm + W.ctx_hi, *const ctx_lo = sm + W.ctx_lo;
    int32_t *const cbuf = reinterpret_cast<int32_t *>(sm + W.c);
    int32_t *const logits = reinterpret_cast<int32_t *>(sm + W.logits);
    int16_t *const scal = reinterpret_cast<int16_t *>(sm + W.scal);
    const int T = a.T, HS = M.h_stride, AS = M.act_stride, NO = M.n_out;
    const bool has_lstm = HS > 0;
    const uint2 *frag = a.frag;

    for (int tile = blockIdx.x; tile * 16 < a.ns; tile += gridDim.x) {
        const int sb = a.s0 + tile * 16;
        const int nvalid = min(16, a.ns - tile * 16);
        /* ---- load the state of the 16 streams: context rows 0..5 of the ring, h buffer 0, c, scalars ---- */
        for (int idx = tid; idx < 16 * 240; idx += MMA_THREADS) {
            const int r = idx / 240, i = idx - r * 240;
            const int v = (r < nvalid) ? (int)a.st.ctx[(long long)(sb + r) * 240 + i] : 0;
            ctx_hi[r * MMA_PC + i] = (uint8_t)(v >> 8);
            ctx_lo[r * MMA_PC + i] = (uint8_t)v;
        }
        for (int idx = tid; idx < 16 * HS; idx += MMA_THREADS) {
            const int r = idx / HS, i = idx - r * HS;
            const int v = (r < nvalid) ? (int)a.st.h[(long long)(sb + r) * HS + i] : 0;
            sm[W.h_hi[0] + r * M.pa + i] = (uint8_t)(v >> 8);
            sm[W.h_lo[0] + r * M.pa + i] = (uint8_t)v;
            cbuf[r * M.hc + i] = (r < nvalid) ? a.st.c[(long long)(sb + r) * HS + i] : 0;
        }
        for (int idx = tid; idx < 16 * SC_N; idx += MMA_THREADS) {
            const int r = idx / SC_N;
            scal[idx] = (r < nvalid) ? a.st.scal[(long long)sb * SC_N + idx] : (int16_t)((idx - r * SC_N) == SC_SLIDES ? 1 : 0);
        }
        int cur = 0, base = 0;                    /* h buffer in use; first row of the context window */
        __syncthreads();
        /* all streams of a batch are reset together and advance in lock step: `slides` is uniform (nn_speech.c:62,125) */
        int slides = scal[SC_SLIDES];

        /* thread -> (stream sr, 5 features from f0) for the standardise step */
        const int sr = tid >> 3, f0 = (tid & 7) * 5;
        const bool srv = sr < nvalid;
        const int32_t *lmrow = a.logmel + ((long long)(sb + (srv ? sr : 0)) * T) * NNSP_B200_NMEL + f0;
        int nxt[5];
#pragma unroll
        for (int j = 0; j < 5; j++) nxt[j] = (srv && !a.raw_ctx) ? __ldg(lmrow + j) : 0;

        for (int t = 0; t < T; t++) {
            /* FeatureClass_execute tail (feature_module.c:54-73): the window slides one ring row */
            int newrow = 0;
            if (a.raw_ctx) {
                /* nnsp_b200_net_eval: rows 0..5 of the ring are the network input as loaded */
            } else if (base == MMA_RING_ROWS - 6) {
                /* rows base+1 .. base+5 back to rows 0..4 (200 bytes per plane row, 50 words, 32 plane rows) */
                __syncthreads();                  /* the previous frame's new row may have been written without a barrier */
                for (int idx = tid; idx < 1600; idx += MMA_THREADS) {
                    const int prow = idx / 50, w = idx - prow * 50;
                    uint32_t *rowp = reinterpret_cast<uint32_t *>(((prow & 1) ? ctx_lo : ctx_hi) + (prow >> 1) * MMA_PC);
                    rowp[w] = rowp[w + 10 * (base + 1)];
                }
                newrow = 5;
                base = 0;
                __syncthreads();
            } else {
                newrow = base + 6;
                base += 1;
            }
            int lm[5];
#pragma unroll
            for (int j = 0; j < 5; j++) lm[j] = nxt[j];
            if (t + 1 < T && !a.raw_ctx) {
#pragma unroll
                for (int j = 0; j < 5; j++) nxt[j] = srv ? __ldg(lmrow + (long long)(t + 1) * NNSP_B200_NMEL + j) : 0;
            }
            if (!a.raw_ctx) {
                const long long ft = (long long)(sb + sr) * T + t;
                const int o = sr * MMA_PC + newrow * 40 + f0;
#pragma unroll
                for (int i = 0; i < 5; i++) {
                    const int v = standardise(lm[i], M.mean[f0 + i], M.stdR[f0 + i], M.feat_rshift);
                    ctx_hi[o + i] = (uint8_t)(v >> 8);
                    ctx_lo[o + i] = (uint8_t)v;
                    if (a.taps.feat && srv) a.taps.feat[ft * 40 + f0 + i] = (int16_t)v;
                    if (a.taps.logmel && srv) a.taps.logmel[ft * 40 + f0 + i] = lm[i];
                }
            }
            const bool ran = (slides == 1);                                                          /* nn_speech.c:84 */
            if (ran) {
                __syncthreads();
                const long long ft0 = (long long)sb * T + t;
                mma_forward(M, frag, bias32, lut, sm, W, 40 * base, cur, warp, MMA_WARPS, lane, nvalid,
                            a.taps.act ? a.taps.act + ft0 * AS : nullptr, (long long)T * AS,
                            a.taps.logits ? a.taps.logits + ft0 * NO : nullptr, (long long)T * NO);
                if (has_lstm) cur ^= 1;
                if (tid < nvalid) {
                    int16_t *sc = scal + tid * SC_N;
                    const int32_t *lg = logits + tid * M.no;
                    if (M.nn_id == NNSP_B200_ID_S2I) post_s2i(sc, lg, a.th_count);                  /* nn_speech.c:97-119 */
                    else post_binary(sc, lg, a.thresh_prob, a.th_count);
                }
            } else {
                if (a.taps.act) for (int i = tid; i < nvalid * AS; i += MMA_THREADS) { const int r = i / AS; a.taps.act[((long long)(sb + r) * T + t) * AS + (i - r * AS)] = 0; }
                if (a.taps.logits) for (int i = tid; i < nvalid * NO; i += MMA_THREADS) { const int r = i / NO; a.taps.logits[((long long)(sb + r) * T + t) * NO + (i - r * NO)] = 0; }
            }
            slides = (slides + 1) % 2;                                                               /* nn_speech.c:125 */
            if (tid < nvalid) {
                int16_t *sc = scal + tid * SC_N;
                sc[SC_SLIDES] = (int16_t)slides;
                if (a.results) {
                    nnsp_b200_result r;
                    r.trigger = sc[SC_TRIGGER];
                    r.outputs[0] = sc[SC_OUT0]; r.outputs[1] = sc[SC_OUT0 + 1]; r.outputs[2] = sc[SC_OUT0 + 2];
                    a.results[(long long)(sb + tid) * T + t] = r;
                }
            }
            if (a.taps.hstate || a.taps.cstate || a.taps.post) {          /* debug taps only: make the frame's state visible */
                __syncthreads();
                const uint8_t *hh = sm + (cur ? W.h_hi[1] : W.h_hi[0]), *hl = sm + (cur ? W.h_lo[1] : W.h_lo[0]);
                for (int i = tid; i < nvalid * HS; i += MMA_THREADS) {
                    const int r = i / HS, u = i - r * HS;
                    const long long o = ((long long)(sb + r) * T + t) * HS + u;
                    if (a.taps.hstate) a.taps.hstate[o] = (int16_t)(((int)(int8_t)hh[r * M.pa + u] << 8) | hl[r * M.pa + u]);
                    if (a.taps.cstate) a.taps.cstate[o] = cbuf[r * M.hc + u];
                }
                if (a.taps.post)
                    for (int i = tid; i < nvalid * SC_N; i += MMA_THREADS) {
                        const int r = i / SC_N, k = i - r * SC_N;
                        int16_t v = scal[i];
                        if (k == SC_RAN) v = ran ? 1 : 0;
                        if (k == SC_STAGE) v = (int16_t)M.nn_id;
                        a.taps.post[((long long)(sb + r) * T + t) * SC_N + k] = v;
                    }
                __syncthreads();
            }
        }
        /* ---- store the state ---- */
        __syncthreads();
        {
            const uint8_t *hh = sm + (cur ? W.h_hi[1] : W.h_hi[0]), *hl = sm + (cur ? W.h_lo[1] : W.h_lo[0]);
            for (int idx = tid; idx < nvalid * 240; idx += MMA_THREADS) {
                const int r = idx / 240, i = idx - r * 240, o = r * MMA_PC + 40 * base + i;
                a.st.ctx[(long long)(sb + r) * 240 + i] = (int16_t)(((int)(int8_t)ctx_hi[o] << 8) | ctx_lo[o]);
            }
            for (int idx = tid; idx < nvalid * HS; idx += MMA_THREADS) {
                const int r = idx / HS, i = idx - r * HS;
                a.st.h[(long long)(sb + r) * HS + i] = (int16_t)(((int)(int8_t)hh[r * M.pa + i] << 8) | hl[r * M.pa + i]);
                a.st.c[(long long)(sb + r) * HS + i] = cbuf[r * M.hc + i];
            }
            for (int idx = tid; idx < nvalid * SC_N; idx += MMA_THREADS) a.st.scal[(long long)sb * SC_N + idx] = scal[idx];
        }
        __syncthreads();
    }
}

/* ---- host: fragment packing -------------------------------------------------------------------- */
static uint32_t pack4(const int8_t *w, int row, int rows, int cols, int k0)
{
    uint32_t v = 0;
    for (int i = 0; i < 4; i++) {
        const int k = k0 + i;
        const uint8_t b = (row >= 0 && row < rows && k < cols) ? (uint8_t)w[(size_t)row * cols + k] : 0;
        v |= (uint32_t)b << (8 * i);
    }
    return v;
}
/* B fragment of mma.m16n8k32 for the 8 rows `rowbase + g` (g = lane / 4), k-step ks */
static void pack_tile(uint2 *dst, const int8_t *w, int nrows_total, int cols, int rowbase, int row_limit, int ks)
{
    for (int lane = 0; lane < 32; lane++) {
        const int g = lane >> 2, q = lane & 3;
        const int row = (g < row_limit) ? rowbase + g : -1;
        dst[lane].x = pack4(w, row, nrows_total, cols, 32 * ks + 4 * q);
        dst[lane].y = pack4(w, row, nrows_total, cols, 32 * ks + 16 + 4 * q);
    }
}

/* Layer 0 in the operand layout of seg0_tc5_kernel (nnsp_tc5.cuh), when the layer qualifies: fc 240 -> rows <= 80, tanh,
 * exact 32-bit finish, an LSTM behind it. Eight K = 32 instructions per byte plane; instruction j, 16-byte chunk c2, byte bb
 * is weight (frame f of the window, feature k):
 *   j 0..3: f = 2 c2 + (j >> 1), k = 16 (j & 1) + bb      j 4, 5: f = 4 + (j - 4), k = 16 c2 + bb
 *   j 6:    f = 2 c2 + (bb >> 3), k = 32 + (bb & 7)        j 7:    c2 = 0: f = 4 + (bb >> 3), k = 32 + (bb & 7); c2 = 1: zero
 * stored K-major without swizzle: [j][unit / 8][c2][unit % 8][16]. The biases follow, shifted as for the IMMA finish. */
static int upload_layer0_tc5(const nnsp_b200_model *m, MmaDeviceModel *out)
{
    const MmaModel *D = out->h;
    const nnsp_layer &L = m->layer[0];
    const MmaLayer &G = D->layer[0];
    if (m->numlayers < 2 || L.type != NNSP_LAYER_FC || m->layer[1].type != NNSP_LAYER_LSTM || L.cols != 240 || L.act != NNSP_ACT_TANH ||
        !G.fast || L.rows > 80)
        return NNSP_B200_OK;
    const int np = (L.rows + 15) & ~15;
    const size_t wbytes = (size_t)8 * np * 32, total = wbytes + (size_t)np * 4;
    uint8_t *img = (uint8_t *)calloc(1, total);
    if (!img) return NNSP_B200_ERR_NOMEM;
    for (int n = 0; n < L.rows; n++) {
        for (int j = 0; j < 8; j++)
            for (int c2 = 0; c2 < 2; c2++)
                for (int bb = 0; bb < 16; bb++) {
                    int f, k;
                    if (j < 4) { f = 2 * c2 + (j >> 1); k = 16 * (j & 1) + bb; }
                    else if (j < 6) { f = 4 + (j - 4); k = 16 * c2 + bb; }
                    else if (j == 6) { f = 2 * c2 + (bb >> 3); k = 32 + (bb & 7); }
                    else { if (c2) continue; f = 4 + (bb >> 3); k = 32 + (bb & 7); }
                    img[(size_t)j * np * 32 + (size_t)(n >> 3) * 256 + c2 * 128 + (n & 7) * 16 + bb] = (uint8_t)L.w[(size_t)n * 240 + f * 40 + k];
                }
        const int32_t b = (int32_t)((uint32_t)(int32_t)L.bias[n] << G.sh_bias);
        memcpy(img + wbytes + (size_t)n * 4, &b, 4);
    }
    cudaError_t e = cudaMalloc(&out->tc5, total);
    if (e == cudaSuccess) e = cudaMemcpy(out->tc5, img, total, cudaMemcpyHostToDevice);
    free(img);
    if (e != cudaSuccess) { nnsp_set_error("tcgen05 layer image upload failed: %s", cudaGetErrorString(e)); return NNSP_B200_ERR_CUDA; }
    out->tc5_np = np;
    return NNSP_B200_OK;
}

int upload_model_mma(const nnsp_b200_model *m, MmaDeviceModel *out)
{
    MmaModel *D = (MmaModel *)calloc(1, sizeof(MmaModel));
    if (!D) return NNSP_B200_ERR_NOMEM;
    out->h = D;
    D->nn_id = m->nn_id; D->numlayers = m->numlayers; D->n_out = m->size_layer[m->numlayers];
    D->feat_rshift = 30 - m->layer[0].qi;
    memcpy(D->mean, m->mean, sizeof D->mean);
    memcpy(D->stdR, m->stdR, sizeof D->stdR);
    for (int i = 0; i < 40; i++) {
        int64_t t = ((int64_t)-147963 - (int64_t)m->mean[i]) * (int64_t)m->stdR[i];
        t >>= D->feat_rshift;
        t = t > 32767 ? 32767 : (t < -32768 ? -32768 : t);
        D->silence[i] = (int16_t)t;
    }
    long long foff = 0; int boff = 0, width = 32, ho = 0;
    for (int i = 0; i < m->numlayers; i++) {
        const nnsp_layer &L = m->layer[i];
        MmaLayer &G = D->layer[i];
        const bool is_lstm = (L.type == NNSP_LAYER_LSTM);
        G.type = L.type; G.act = L.act; G.rows = L.rows; G.cols = L.cols; G.acc32 = L.acc32;
        G.kt = (L.cols + 31) / 32; G.ktr = is_lstm ? (L.rows + 31) / 32 : 0;
        G.nt = (L.rows + 7) / 8;
        const int qi_out = is_lstm ? L.qi_next : L.qi;
        const int qs = (qi_out + L.qk) > 15 ? (qi_out + L.qk) : 15;
        G.sh_x = is_lstm ? (L.qi_next - L.qi) : 0;
        G.sh_bias = qs - L.qb;
        G.sh_out = 15 - qs;
        if (L.cols > 480 || L.rows > NNSP_B200_MAX_WIDTH || (i == 0 && L.cols != 240) || (i > 0 && L.cols > NNSP_B200_MAX_WIDTH)) return NNSP_B200_ERR_UNSUPPORTED;
        if (is_lstm && (ho & 3)) return NNSP_B200_ERR_UNSUPPORTED;
        /* exact 32-bit finish: no clamp of the reference can fire and every shift is an arithmetic right shift */
        const double macs = (double)(L.cols + (is_lstm ? L.rows : 0)) * 128.0 * 32768.0;
        const bool shifts_ok = (G.sh_x == 0 && G.sh_bias >= 0 && G.sh_bias <= 15 && G.sh_out <= 0 && G.sh_out >= -28);   /* -28: the relu6 finish shifts by 3 more */
        G.fast = shifts_ok && (L.acc32 || macs + 32768.0 * (double)(1 << G.sh_bias) < 2147483647.0);
        G.w_off = (int)foff;
        foff += (long long)(is_lstm ? 4 : 1) * G.nt * G.kt * 32;
        G.wh_off = (int)foff;
        foff += (long long)(is_lstm ? 4 : 0) * G.nt * G.ktr * 32;
        boff = (boff + 1) & ~1;                            /* even: the fc finish reads bias pairs with one 64-bit load */
        G.bias_off = boff; boff += (is_lstm ? 4 : 1) * L.rows;
        if (i < m->numlayers - 1) D->act_stride += L.rows;
        if (i > 0 && G.kt * 32 > width) width = G.kt * 32;
        if (G.nt * 8 > width) width = G.nt * 8;
        if (is_lstm) { if (ho + G.ktr * 32 > width) width = ho + G.ktr * 32; ho += L.rows; D->h_stride += L.rows; }
    }
    if (D->h_stride > NNSP_B200_MAX_WIDTH) return NNSP_B200_ERR_UNSUPPORTED;
    int P = (width + 3) / 4;
    while ((P & 7) != 4) P++;                              /* pitch/4 = 4 (mod 8): conflict-free A-fragment loads */
    D->pa = 4 * P;
    D->hc = (D->h_stride + 7) & ~7; if (D->hc == 0) D->hc = 8;
    D->no = (D->n_out + 7) & ~7;
    D->frag_count = (int)((foff + 1) & ~1LL);
    D->bias_count = (boff + 8 + 3) & ~3;                   /* 8 of slack: padded units read past their layer's rows */
    D->warp_bytes = (2 * 16 * MMA_PC + 8 * 16 * D->pa + 16 * D->hc * 4 + 16 * D->no * 4 + 16 * SC_N * 2 + 15) & ~15;
    uint2 *hf = (uint2 *)calloc((size_t)D->frag_count + 2, sizeof(uint2));
    int32_t *hb = (int32_t *)calloc((size_t)D->bias_count + 4, sizeof(int32_t));
    if (!hf || !hb) { free(hf); free(hb); return NNSP_B200_ERR_NOMEM; }
    for (int i = 0; i < m->numlayers; i++) {
        const nnsp_layer &L = m->layer[i];
        const MmaLayer &G = D->layer[i];
        if (L.type == NNSP_LAYER_LSTM) {
            const int H = L.rows;
            for (int grp = 0; grp < G.nt; grp++)
                for (int gt = 0; gt < 4; gt++) {
                    const int lim = (H - 8 * grp) < 8 ? (H - 8 * grp) : 8;
                    for (int ks = 0; ks < G.kt; ks++)
                        pack_tile(hf + G.w_off + ((size_t)(grp * 4 + gt) * G.kt + ks) * 32, L.w, 4 * H, L.cols, gt * H + 8 * grp, lim, ks);
                    for (int ks = 0; ks < G.ktr; ks++)
                        pack_tile(hf + G.wh_off + ((size_t)(grp * 4 + gt) * G.ktr + ks) * 32, L.wrec, 4 * H, H, gt * H + 8 * grp, lim, ks);
                }
            for (int n = 0; n < 4 * H; n++) hb[G.bias_off + n] = G.fast ? (int32_t)((uint32_t)(int32_t)L.bias[n] << G.sh_bias) : (int32_t)L.bias[n];
        } else {
            for (int nt = 0; nt < G.nt; nt++) {
                const int lim = (L.rows - 8 * nt) < 8 ? (L.rows - 8 * nt) : 8;
                for (int ks = 0; ks < G.kt; ks++)
                    pack_tile(hf + G.w_off + ((size_t)nt * G.kt + ks) * 32, L.w, L.rows, L.cols, 8 * nt, lim, ks);
            }
            for (int n = 0; n < L.rows; n++) hb[G.bias_off + n] = G.fast ? (int32_t)((uint32_t)(int32_t)L.bias[n] << G.sh_bias) : (int32_t)L.bias[n];
        }
    }
    auto a16 = [](size_t v) { return (v + 15) & ~(size_t)15; };
    out->off_bias = 0;                                       /* bias32 | tanh LUT | model | one 16-stream tile */
    out->off_lut = (int)a16((size_t)D->bias_count * 4);
    out->off_model = (int)a16(out->off_lut + 384 * 2);
    out->off_warps = (int)((out->off_model + sizeof(MmaModel) + 127) & ~(size_t)127);
    out->smem_base = out->off_warps;
    out->smem_warp = D->warp_bytes;
    cudaError_t e1 = cudaSuccess, e2 = cudaSuccess, e3 = cudaSuccess;
    if (out->smem_base + out->smem_warp > 227 * 1024) { free(hf); free(hb); return NNSP_B200_ERR_UNSUPPORTED; }
    e1 = cudaMalloc(&out->frag, (size_t)D->frag_count * 8);
    e2 = cudaMalloc(&out->bias32, (size_t)D->bias_count * 4);
    e3 = cudaMalloc(&out->d, sizeof(MmaModel));
    if (e1 == cudaSuccess && e2 == cudaSuccess && e3 == cudaSuccess) {
        e1 = cudaMemcpy(out->frag, hf, (size_t)D->frag_count * 8, cudaMemcpyHostToDevice);
        e2 = cudaMemcpy(out->bias32, hb, (size_t)D->bias_count * 4, cudaMemcpyHostToDevice);
        e3 = cudaMemcpy(out->d, D, sizeof(MmaModel), cudaMemcpyHostToDevice);
    }
    free(hf); free(hb);
    if (e1 != cudaSuccess || e2 != cudaSuccess || e3 != cudaSuccess) {
        nnsp_set_error("IMMA model upload failed: %s", cudaGetErrorString(e1 != cudaSuccess ? e1 : (e2 != cudaSuccess ? e2 : e3)));
        return NNSP_B200_ERR_CUDA;
    }
    return upload_layer0_tc5(m, out);
}

void free_model_mma(MmaDeviceModel *mm)
{
    if (mm->frag) cudaFree(mm->frag);
    if (mm->bias32) cudaFree(mm->bias32);
    if (mm->d) cudaFree(mm->d);
    if (mm->tc5) cudaFree(mm->tc5);
    free(mm->h);
    mm->tc5 = nullptr; mm->tc5_np = 0;
    mm->frag = nullptr; mm->bias32 = nullptr; mm->d = nullptr; mm->h = nullptr;
}

int launch_nn_mma(const MmaDeviceModel &mm, const NNLaunch &l, int device, cudaStream_t st)
{
    const size_t smem = mm.smem_base + mm.smem_warp;
    static bool attr_done[64] = { false };
    static std::mutex attr_mu;                         /* host threads may drive separate handles on one device */
    std::unique_lock<std::mutex> attr_lk(attr_mu);
    if (!attr_done[device]) {
        NNSP_CUDA(cudaFuncSetAttribute(nn_mma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
        attr_done[device] = true;
    }
    attr_lk.unlock();
    MmaArgs a{};
    a.model = mm.d; a.frag = (const uint2 *)mm.frag; a.bias32 = mm.bias32; a.tables = l.tables; a.st = l.st;
    a.logmel = l.logmel; a.s0 = l.s0; a.ns = l.ns; a.T = l.T; a.results = l.results; a.taps = l.taps;
    a.thresh_prob = l.thresh_prob; a.th_count = l.th_count; a.raw_ctx = l.raw_ctx;
    int blocks = (l.ns + 15) / 16;
    /* leave L1 room for the weight image: the B fragments are read through the read-only L1 path */
    const size_t frag_bytes = (size_t)mm.h->frag_count * 8;
    int per_sm = (int)((227 * 1024 - (frag_bytes < 160 * 1024 ? frag_bytes : 160 * 1024)) / smem);
    if (per_sm < 1) per_sm = 1;
    if (per_sm > 4) per_sm = 4;                              /* __launch_bounds__(128, 4): <= 128 registers */
    const int cap = sm_count(device) * per_sm;
    if (blocks > cap) blocks = cap;
    nn_mma_kernel<<<blocks, MMA_THREADS, smem, st>>>(a, mm.off_lut, mm.off_model, mm.off_warps, (int)smem);
    NNSP_LAUNCH_CHECK();
    return NNSP_B200_OK;
}

}  // namespace nnsp
