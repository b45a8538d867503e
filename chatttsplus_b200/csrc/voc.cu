// Vocoder: DVAE decode (dvae.py:254-291) + Vocos decode (pip `vocos`: VocosBackbone, ISTFTHead) for a whole batch
// of variable-length utterances in one pass (the reference loops utterance by utterance at batch 1,
// chattts_plus_pipeline.py:298-304).
//
// Layout: all utterances are laid out on one time axis of "mel frames" (2 per code frame), channels-last
// [rows][C], separated by GAP zero rows so that every convolution sees exactly the zero padding the reference's
// per-utterance Conv1d(padding=...) sees.  Every convolution is then a tcgen05 GEMM:
//   * k=3 / k=7 dense convs: im2col is free — row r's window is the contiguous span starting at row r-1 (r-3), so the
//     A operand is a TMA tensor map with row pitch C and row length 3C (7C): overlapping rows;
//   * 1x1 convs / ConvNeXt pointwise layers: plain GEMMs with fused bias / GELU / layer-scale / residual epilogues.
// Depthwise conv + LayerNorm, the ISTFT (shared-memory inverse FFT) and overlap-add are CUDA-core kernels.
#include "gemm.cuh"
#include "../../include/ctp.h"

#include <algorithm>
#include <math.h>
#include <string.h>
#include <vector>

using namespace ctp;

namespace {

constexpr int GAP = 8;       // zero rows before the first, between, and after the last utterance (>= 6 = k7 dil2 reach)
constexpr int MEL_PAD = 104; // mel channels padded so that the im2col row pitch (208 B) is a multiple of 16 B

struct UttTable {
    int n_utt;
    const int* row0;     // [n_utt] first row of utterance i
    const int* nrows;    // [n_utt] mel frames of utterance i
};

// ---- input formatting -----------------------------------------------------------------------------------
// hiddens fp32 [n][2*idim] -> X0 fp16 rows (frame 2t+j = hid[t][j*idim .. (j+1)*idim)): dvae.py:277-283 is a pure view
// in channels-last layout.  One CTA per row of the padded time axis; invalid rows are zeroed.
__global__ void k_voc_input_hidden(const float* __restrict__ hid, __half* __restrict__ x0, const int* __restrict__ row_src,
                                   int idim) {
    const int r = blockIdx.x;
    const int src = row_src[r];  // index into the flat [sum 2 n_i][idim] view, or -1
    __half* o = x0 + (long long)r * idim;
    for (int c = threadIdx.x; c < idim; c += blockDim.x) o[c] = (src >= 0) ? __float2half_rn(hid[(long long)src * idim + c]) : __half(0);
}

// GFSQ embed (dvae.py:84-94 -> vector_quantize_pytorch GroupedResidualFSQ.get_output_from_indices, levels 5^4,
// G groups x R residual levels): codes int32 [n][G*R] -> feat [n][G*gd]; then the same 2-frame interleave.
__global__ void k_voc_input_codes(const int* __restrict__ codes, __half* __restrict__ x0, const int* __restrict__ row_src,
                                  const float* __restrict__ proj_w, const float* __restrict__ proj_b, int idim, int G, int R) {
    const int r = blockIdx.x;
    const int src = row_src[r];  // flat frame index 2t+j or -1
    __half* o = x0 + (long long)r * idim;
    if (src < 0) {
        for (int c = threadIdx.x; c < idim; c += blockDim.x) o[c] = __half(0);
        return;
    }
    const int t = src >> 1, j = src & 1;
    const int dim = 2 * idim, gd = dim / G;
    __shared__ float code[8][4];
    if (threadIdx.x < G * 4) {
        const int g = threadIdx.x >> 2, d = threadIdx.x & 3;
        float acc = 0.f, scale = 1.f;
        int basis = 1;
        for (int i = 0; i < d; ++i) basis *= 5;
        for (int rr = 0; rr < R; ++rr) {
            const int idx = codes[(long long)t * G * R + g * R + rr];
            const int digit = (idx / basis) % 5;
            acc += ((float)digit - 2.0f) * 0.5f * scale;  // (digit - half) / half, half = 2
            scale *= 0.25f;                               // (levels - 1)^-1
        }
        code[g][d] = acc;
    }
    __syncthreads();
    for (int c = threadIdx.x; c < idim; c += blockDim.x) {
        const int ch = j * idim + c;  // channel in the [dim] feature vector
        const int g = ch / gd, cc = ch % gd;
        const float* w = proj_w + ((long long)g * gd + cc) * 4;
        const float v = proj_b[g * gd + cc] + w[0] * code[g][0] + w[1] * code[g][1] + w[2] * code[g][2] + w[3] * code[g][3];
        o[c] = __float2half_rn(v);
    }
}

// ---- depthwise conv (k=7, dilation d) + LayerNorm(eps 1e-6) -> fp16 GEMM operand  (dvae.py:48-54) ---------------
template <int C>
__global__ void __launch_bounds__(128) k_dwconv_ln(const float* __restrict__ x, __half* __restrict__ y, const unsigned char* __restrict__ valid,
                                                    const float* __restrict__ dw_w, const float* __restrict__ dw_b,
                                                    const float* __restrict__ ln_w, const float* __restrict__ ln_b, int dil) {
    constexpr int PER = C / 128;
    const int r = blockIdx.x;
    __half* o = y + (long long)r * C;
    if (!valid[r]) {
#pragma unroll
        for (int i = 0; i < PER; ++i) o[threadIdx.x + i * 128] = __half(0);
        return;
    }
    __shared__ float red[4];
    float v[PER];
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < PER; ++i) {
        const int c = threadIdx.x + i * 128;
        float acc = dw_b[c];
#pragma unroll
        for (int k = 0; k < 7; ++k) acc += dw_w[c * 7 + k] * x[(long long)(r + (k - 3) * dil) * C + c];
        v[i] = acc;
        s += acc;
    }
    s = warp_sum(s);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
    __syncthreads();
    const float mean = (red[0] + red[1] + red[2] + red[3]) / (float)C;
    __syncthreads();
    float q = 0.f;
#pragma unroll
    for (int i = 0; i < PER; ++i) { const float d = v[i] - mean; q += d * d; }
    q = warp_sum(q);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = q;
    __syncthreads();
    const float rstd = rsqrtf((red[0] + red[1] + red[2] + red[3]) / (float)C + 1e-6f);
#pragma unroll
    for (int i = 0; i < PER; ++i) {
        const int c = threadIdx.x + i * 128;
        o[c] = __float2half_rn((v[i] - mean) * rstd * ln_w[c] + ln_b[c]);
    }
}

// Same operator, R consecutive rows per CTA with a sliding tap window in registers: a row of the residual stream is pulled from L2
// (R + 6 dil) / R times instead of 7 times (one CTA per row re-read every row for each of its seven taps: 14 KB of L2 traffic per
// 2 KB row, 60 us per 32000 rows against a 15 us HBM floor), the 7 tap weights per channel are fetched once per CTA, and the two
// LayerNorm reductions are shared by four rows.  Rows of one residue class mod dil form one sliding chain.  The per-row arithmetic
// (tap order, reduction order) is that of k_dwconv_ln, so the results are bit-identical.
template <int C, int R>
__global__ void __launch_bounds__(128) k_dwconv_ln_tile(const float* __restrict__ x, __half* __restrict__ y, const unsigned char* __restrict__ valid,
                                                         const float* __restrict__ dw_w, const float* __restrict__ dw_b,
                                                         const float* __restrict__ ln_w, const float* __restrict__ ln_b, int dil, int rows) {
    constexpr int PER = C / 128;
    const int r0 = blockIdx.x * R;
    const int warp = threadIdx.x >> 5;
    __shared__ float red[2][4][4];   // [mean | variance pass][row of the batch][warp]
    float w[PER][7], bias[PER], lw[PER], lb[PER];
#pragma unroll
    for (int i = 0; i < PER; ++i) {
        const int c = threadIdx.x + i * 128;
#pragma unroll
        for (int k = 0; k < 7; ++k) w[i][k] = dw_w[c * 7 + k];
        bias[i] = dw_b[c]; lw[i] = ln_w[c]; lb[i] = ln_b[c];
    }
    const long long lim = (long long)rows + GAP;   // rows [rows, rows + GAP) are the zero guard; nothing is addressable past it
    for (int p = 0; p < dil; ++p) {
        float win[PER][7];
        const int rf = r0 + p;
#pragma unroll
        for (int i = 0; i < PER; ++i) {
#pragma unroll
            for (int k = 0; k < 6; ++k) {
                const long long q = (long long)rf + (k - 3) * dil;
                win[i][k] = q < lim ? x[q * C + threadIdx.x + i * 128] : 0.f;
            }
        }
        // the newest tap row of each of the next four output rows is requested one round ahead (16 independent loads in flight per
        // thread while the current round computes and reduces); a load issued right where it is consumed left every row exposed to a
        // full L2 / HBM latency (230 us per 32000 rows, measured)
        float nxt[4][PER];
#pragma unroll
        for (int bb = 0; bb < 4; ++bb) {
#pragma unroll
            for (int i = 0; i < PER; ++i) {
                const long long q = (long long)rf + (long long)bb * dil + 3 * dil;
                nxt[bb][i] = q < lim ? x[q * C + threadIdx.x + i * 128] : 0.f;
            }
        }
        for (int j0 = 0; j0 < R / dil; j0 += 4) {   // four rows of this residue class per reduction round
            float cur[4][PER];
#pragma unroll
            for (int bb = 0; bb < 4; ++bb) {
#pragma unroll
                for (int i = 0; i < PER; ++i) cur[bb][i] = nxt[bb][i];
            }
            if (j0 + 4 < R / dil) {
#pragma unroll
                for (int bb = 0; bb < 4; ++bb) {
#pragma unroll
                    for (int i = 0; i < PER; ++i) {
                        const long long q = (long long)rf + (long long)(j0 + 4 + bb) * dil + 3 * dil;
                        nxt[bb][i] = q < lim ? x[q * C + threadIdx.x + i * 128] : 0.f;
                    }
                }
            }
            float v[4][PER], s4[4];
#pragma unroll
            for (int bb = 0; bb < 4; ++bb) {
                float sacc = 0.f;
#pragma unroll
                for (int i = 0; i < PER; ++i) {
                    win[i][6] = cur[bb][i];
                    float acc = bias[i];
#pragma unroll
                    for (int k = 0; k < 7; ++k) acc += w[i][k] * win[i][k];
#pragma unroll
                    for (int k = 0; k < 6; ++k) win[i][k] = win[i][k + 1];
                    v[bb][i] = acc;
                    sacc += acc;
                }
                s4[bb] = warp_sum(sacc);
            }
            if ((threadIdx.x & 31) == 0) {
#pragma unroll
                for (int bb = 0; bb < 4; ++bb) red[0][bb][warp] = s4[bb];
            }
            __syncthreads();
            float mean[4];
#pragma unroll
            for (int bb = 0; bb < 4; ++bb) {
                mean[bb] = (red[0][bb][0] + red[0][bb][1] + red[0][bb][2] + red[0][bb][3]) / (float)C;
                float q = 0.f;
#pragma unroll
                for (int i = 0; i < PER; ++i) { const float d = v[bb][i] - mean[bb]; q += d * d; }
                s4[bb] = warp_sum(q);
            }
            if ((threadIdx.x & 31) == 0) {
#pragma unroll
                for (int bb = 0; bb < 4; ++bb) red[1][bb][warp] = s4[bb];
            }
            __syncthreads();
#pragma unroll
            for (int bb = 0; bb < 4; ++bb) {
                const int r = rf + (j0 + bb) * dil;
                if (r >= rows) continue;
                const float rstd = rsqrtf((red[1][bb][0] + red[1][bb][1] + red[1][bb][2] + red[1][bb][3]) / (float)C + 1e-6f);
                const bool ok = valid[r] != 0;
                __half* o = y + (long long)r * C;
#pragma unroll
                for (int i = 0; i < PER; ++i)
                    o[threadIdx.x + i * 128] = ok ? __float2half_rn((v[bb][i] - mean[bb]) * rstd * lw[i] + lb[i]) : __half(0);
            }
            // (the next round's first write to red[0] is ordered behind this round's reads of red[0] by the second barrier, and its
            // writes to red[1] behind these reads of red[1] by its first barrier)
        }
    }
}

// LayerNorm over channels of fp32 rows -> fp32 (in place allowed) and/or fp16
template <int C>
__global__ void __launch_bounds__(128) k_ln_rows(const float* __restrict__ x, float* __restrict__ out32, __half* __restrict__ out16,
                                                  const unsigned char* __restrict__ valid, const float* __restrict__ w,
                                                  const float* __restrict__ b) {
    constexpr int PER = C / 128;
    const int r = blockIdx.x;
    if (!valid[r]) {
#pragma unroll
        for (int i = 0; i < PER; ++i) {
            if (out32) out32[(long long)r * C + threadIdx.x + i * 128] = 0.f;
            if (out16) out16[(long long)r * C + threadIdx.x + i * 128] = __half(0);
        }
        return;
    }
    __shared__ float red[4];
    float v[PER];
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < PER; ++i) { v[i] = x[(long long)r * C + threadIdx.x + i * 128]; s += v[i]; }
    s = warp_sum(s);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
    __syncthreads();
    const float mean = (red[0] + red[1] + red[2] + red[3]) / (float)C;
    __syncthreads();
    float q = 0.f;
#pragma unroll
    for (int i = 0; i < PER; ++i) { const float d = v[i] - mean; q += d * d; }
    q = warp_sum(q);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = q;
    __syncthreads();
    const float rstd = rsqrtf((red[0] + red[1] + red[2] + red[3]) / (float)C + 1e-6f);
#pragma unroll
    for (int i = 0; i < PER; ++i) {
        const int c = threadIdx.x + i * 128;
        const float yv = (v[i] - mean) * rstd * w[c] + b[c];
        if (out32) out32[(long long)r * C + c] = yv;
        if (out16) out16[(long long)r * C + c] = __float2half_rn(yv);
    }
}

// fp32 [rows][C] -> fp16 [rows][Cpad] (extra channels zero); invalid rows zero
__global__ void k_cvt_rows(const float* __restrict__ x, __half* __restrict__ y, const unsigned char* __restrict__ valid, int C, int Cpad) {
    const int r = blockIdx.x;
    const bool ok = valid[r] != 0;
    for (int c = threadIdx.x; c < Cpad; c += blockDim.x)
        y[(long long)r * Cpad + c] = (ok && c < C) ? __float2half_rn(x[(long long)r * C + c]) : __half(0);
}

// Row pitch of the ISTFT head buffer [rows][n_fft + 2 -> 1028]: a multiple of four floats so that the GEMM epilogue's float4 stores
// apply (with the natural pitch of 1026 every second row is only 8-byte aligned and the whole tile fell back to scalar stores: the
// head GEMM ran at 115 TFLOP/s against 740 for the same K at N = 1536)
constexpr int HEAD_PITCH = 1028;

// ---- ISTFT head (vocos ISTFTHead + ISTFT(padding="center") == torch.istft(center=True)) ---------------------------
// One CTA (256 threads) per mel frame: S = exp(mag) clipped at 1e2 times (cos p, sin p), Hermitian extension, 1024-point inverse
// FFT in shared memory as FIVE radix-4 Stockham passes (natural order in and out, no bit reversal; one 4-point butterfly per thread
// and pass, reads at stride 256 are conflict-free), multiply by the synthesis window; store the windowed frame.  (The first version ran
// ten radix-2 passes with two butterflies per thread and a barrier each: 565 us per 32000 frames against a 40 us HBM floor.)
__global__ void __launch_bounds__(256) k_istft_frame(const float* __restrict__ head, float* __restrict__ frames,
                                                      const unsigned char* __restrict__ valid, const float* __restrict__ window,
                                                      const float2* __restrict__ twiddle /*[512] exp(+2 pi i k/1024)*/) {
    constexpr int N = 1024, NB = 513;
    const int r = blockIdx.x;
    if (!valid[r]) return;
    __shared__ float2 buf0[N], buf1[N];
    __shared__ float2 tw[512];
    const float* hp = head + (long long)r * HEAD_PITCH;
    for (int k = threadIdx.x; k < 512; k += 256) tw[k] = twiddle[k];
    for (int k = threadIdx.x; k < N; k += 256) {
        const int kk = (k <= 512) ? k : (N - k);
        float mag = __expf(hp[kk]);
        mag = fminf(mag, 1e2f);
        float sn, cs;
        sincosf(hp[NB + kk], &sn, &cs);
        float re = mag * cs, im = mag * sn;
        if (kk == 0 || kk == 512) im = 0.f;  // c2r ignores the imaginary part of DC / Nyquist
        if (k > 512) im = -im;
        buf0[k] = make_float2(re, im);
    }
    __syncthreads();
    const int j = threadIdx.x;
    float2* in = buf0;
    float2* out = buf1;
#pragma unroll
    for (int s = 0; s < 5; ++s) {
        const int Ns = 1 << (2 * s);          // 1, 4, 16, 64, 256
        const int k = j & (Ns - 1);
        const int step = 256 >> (2 * s);      // twiddle of input q: exp(+2 pi i q k / (4 Ns)) = W1024^(q k step)
        float2 v0 = in[j], v1 = in[j + 256], v2 = in[j + 512], v3 = in[j + 768];
        if (s > 0) {
            const int i1 = k * step, i2 = 2 * k * step, i3 = 3 * k * step;   // < 1024
            float2 w1 = tw[i1 & 511], w2 = tw[i2 & 511], w3 = tw[i3 & 511];
            if (i1 & 512) { w1.x = -w1.x; w1.y = -w1.y; }
            if (i2 & 512) { w2.x = -w2.x; w2.y = -w2.y; }
            if (i3 & 512) { w3.x = -w3.x; w3.y = -w3.y; }
            v1 = make_float2(v1.x * w1.x - v1.y * w1.y, v1.x * w1.y + v1.y * w1.x);
            v2 = make_float2(v2.x * w2.x - v2.y * w2.y, v2.x * w2.y + v2.y * w2.x);
            v3 = make_float2(v3.x * w3.x - v3.y * w3.y, v3.x * w3.y + v3.y * w3.x);
        }
        const float2 t0 = make_float2(v0.x + v2.x, v0.y + v2.y), t1 = make_float2(v0.x - v2.x, v0.y - v2.y);
        const float2 t2 = make_float2(v1.x + v3.x, v1.y + v3.y);
        const float2 t3 = make_float2(-(v1.y - v3.y), v1.x - v3.x);   // +i (v1 - v3): inverse transform
        const int j0 = ((j - k) << 2) + k;
        out[j0] = make_float2(t0.x + t2.x, t0.y + t2.y);
        out[j0 + Ns] = make_float2(t1.x + t3.x, t1.y + t3.y);
        out[j0 + 2 * Ns] = make_float2(t0.x - t2.x, t0.y - t2.y);
        out[j0 + 3 * Ns] = make_float2(t1.x - t3.x, t1.y - t3.y);
        __syncthreads();
        float2* t = in; in = out; out = t;
    }
    float* o = frames + (long long)r * N;
    for (int n = threadIdx.x; n < N; n += 256) o[n] = in[n].x * (1.0f / N) * window[n];
}

// overlap-add with window-envelope normalisation and centre trimming.  grid (ceil(max_len/256), n_utt)
__global__ void k_overlap_add(const float* __restrict__ frames, float* __restrict__ wav, const int* __restrict__ row0,
                              const int* __restrict__ nrows, const long long* __restrict__ wav_off, const float* __restrict__ window,
                              int hop, int n_fft) {
    const int u = blockIdx.y;
    const int T = nrows[u];
    const long long len = (long long)hop * (T - 1);
    const long long n = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (n >= len) return;
    const long long m = n + n_fft / 2;
    int f_hi = (int)(m / hop);
    if (f_hi > T - 1) f_hi = T - 1;
    long long f_lo_ll = (m - n_fft) / hop + 1;
    if (m - n_fft < 0) f_lo_ll = 0;
    int f_lo = (int)f_lo_ll;
    if (f_lo < 0) f_lo = 0;
    float acc = 0.f, env = 0.f;
    for (int f = f_lo; f <= f_hi; ++f) {
        const int k = (int)(m - (long long)f * hop);
        if (k < 0 || k >= n_fft) continue;
        acc += frames[(long long)(row0[u] + f) * n_fft + k];
        const float w = window[k];
        env += w * w;
    }
    wav[wav_off[u] + n] = acc / env;
}

// ---- prompt encoder front end: MelSpectrogramFeatures (dvae.py:171-199 = torchaudio MelSpectrogram(power=1, center=True)) --
// One CTA (256 threads) per frame: 1024 samples around f*hop with reflect padding, times the periodic Hann window, forward
// real FFT in shared memory (radix-2, conjugated twiddles of the ISTFT kernel), |X[k]| for k <= 512, mel = fb^T |X|,
// y = log(max(mel, 1e-5)) / coef  (dvae.py:198,266) stored as the fp16 operand row of downsample_conv.0.
__global__ void __launch_bounds__(256) k_mel_frame(const float* __restrict__ audio, int n_samples, __half* __restrict__ mel16 /* row of frame 0 */,
                                                    const float* __restrict__ window, const float2* __restrict__ twiddle,
                                                    const float* __restrict__ fb /*[513][n_mels]*/, const float* __restrict__ coef,
                                                    int n_mels, int hop) {
    constexpr int N = 1024, NB = 513;
    const int f = blockIdx.x;
    __shared__ float2 a[N];
    __shared__ float mag[NB];
    for (int j = threadIdx.x; j < N; j += 256) {
        long long n = (long long)f * hop - N / 2 + j;
        if (n < 0) n = -n;                                   // reflect (no edge repeat), torch.stft(center=True, pad_mode="reflect")
        if (n >= n_samples) n = 2LL * (n_samples - 1) - n;
        const float v = audio[n] * window[j];
        const int rev = __brev((unsigned)j) >> 22;
        a[rev] = make_float2(v, 0.f);
    }
    __syncthreads();
#pragma unroll 1
    for (int s = 1; s <= 10; ++s) {
        const int half = 1 << (s - 1);
        for (int j = threadIdx.x; j < N / 2; j += 256) {
            const int grp = j >> (s - 1), pos = j & (half - 1);
            const int i0 = (grp << s) + pos, i1 = i0 + half;
            const float2 w0 = twiddle[pos << (10 - s)];
            const float2 w = make_float2(w0.x, -w0.y);       // forward transform: exp(-2 pi i k / N)
            const float2 x1 = a[i1], x0 = a[i0];
            const float2 t = make_float2(w.x * x1.x - w.y * x1.y, w.x * x1.y + w.y * x1.x);
            a[i1] = make_float2(x0.x - t.x, x0.y - t.y);
            a[i0] = make_float2(x0.x + t.x, x0.y + t.y);
        }
        __syncthreads();
    }
    for (int k = threadIdx.x; k < NB; k += 256) mag[k] = sqrtf(a[k].x * a[k].x + a[k].y * a[k].y);
    __syncthreads();
    __half* o = mel16 + (long long)f * MEL_PAD;
    for (int m = threadIdx.x; m < MEL_PAD; m += 256) {
        float y = 0.f;
        if (m < n_mels) {
            float acc = 0.f;
            for (int k = 0; k < NB; ++k) acc += fb[k * n_mels + m] * mag[k];
            y = __logf(fmaxf(acc, 1e-5f)) / coef[m];
        }
        o[m] = __float2half_rn(y);
    }
}

// ---- GFSQ quantiser (dvae.py:98-126 -> vector_quantize_pytorch GroupedResidualFSQ.forward, levels [5,5,5,5], G groups, R = 2) ----
// One CTA per code frame, warp g handles group g: z = project_in(x_g) (4 dots of odim/G), residual = bound(z) = tanh(z) * 2.002,
// then per residual level r: q = rint(tanh(residual / s_r) * 2.002) / 2, digit_j = 2 q_j + 2, index = sum digit_j 5^j,
// residual -= q s_r with s_r = 4^-r.
__global__ void k_gfsq_quantize(const float* __restrict__ x /*[T][odim]*/, int odim, int G, const float* __restrict__ w /*[G][4][odim/G]*/,
                                const float* __restrict__ b /*[G][4]*/, int* __restrict__ ids /*[T][G*2]*/) {
    const int t = blockIdx.x, g = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (g >= G) return;
    const int gd = odim / G;
    const float* xr = x + (long long)t * odim + g * gd;
    float z[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        const float* wr = w + ((long long)g * 4 + j) * gd;
        float acc = 0.f;
        for (int c = lane; c < gd; c += 32) acc += wr[c] * xr[c];
        z[j] = warp_sum(acc) + b[g * 4 + j];
    }
    if (lane == 0) {
        const float half_l = 4.0f * (1.0f + 1e-3f) / 2.0f;   // (levels - 1) * (1 + eps) / 2
        float res[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) res[j] = tanhf(z[j]) * half_l;
        float scale = 1.0f;
        for (int r = 0; r < 2; ++r) {
            int idx = 0, basis = 1;
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const float q = rintf(tanhf(res[j] / scale) * half_l) * 0.5f;   // torch.round = round-half-to-even
                idx += (int)rintf(q * 2.0f + 2.0f) * basis;
                basis *= 5;
                res[j] -= q * scale;
            }
            ids[(long long)t * (G * 2) + g * 2 + r] = idx;
            scale *= 0.25f;
        }
    }
}

struct ConvNextDev {
    ctp_convnext_w w;
};

}  // namespace

struct ctp_voc {
    ctp_voc_cfg cfg{};
    ctp_voc_weights w{};
    std::vector<ctp_convnext_w> dvae_blocks, voc_blocks;
    bool bound = false, has_dvae = false, has_vocos = false, has_encoder = false;
    long long cap_rows = 0;  // capacity in rows (excluding the +-GAP guard rows)
    // activation buffers; every pointer is offset by GAP rows so that rows [-GAP, cap+GAP) are addressable
    __half* x0 = nullptr;     // [rows][idim]
    __half* c1 = nullptr;     // [rows][bn]
    float* xres = nullptr;    // [rows][max(hidden, voc_dim)]
    __half* y16 = nullptr;    // [rows][max(hidden, voc_dim)]
    __half* hm = nullptr;     // [rows][max(4*hidden, voc_inter)]
    __half* o1 = nullptr;     // [rows][odim]
    float* mel32 = nullptr;   // [rows][n_mels]
    __half* mel16 = nullptr;  // [rows][MEL_PAD]
    int gelu_mode = 2;        // GEMM epilogue GELU: 2 = one-MUFU erf (ctp_common.cuh), 1 = Abramowitz-Stegun form (CTP_GELU=as)
    bool dw_tile = true;      // k_dwconv_ln_tile (16 rows per CTA); CTP_DWCONV=row keeps one CTA per row
    float* head = nullptr;    // [rows][n_fft + 2, pitch HEAD_PITCH]
    float* frames = nullptr;  // [rows][n_fft]
    unsigned char* valid = nullptr;  // [rows]
    int* row_src = nullptr;   // [rows]
    int* d_row0 = nullptr; int* d_nrows = nullptr; long long* d_wavoff = nullptr;  // [max_utt]
    float2* twiddle = nullptr;
    int max_utt = 4096;
    std::vector<void*> allocs;
};

template <typename T>
static int voc_alloc(ctp_voc* h, T** p, long long rows, long long cols) {
    const size_t bytes = (size_t)(rows + 2 * GAP) * cols * sizeof(T);
    void* raw = nullptr;
    cudaError_t e = cudaMalloc(&raw, bytes);
    if (e != cudaSuccess) { ctp_set_error("voc workspace cudaMalloc(%zu): %s", bytes, cudaGetErrorString(e)); return CTP_ERR_CUDA; }
    cudaMemset(raw, 0, bytes);
    h->allocs.push_back(raw);
    *p = reinterpret_cast<T*>(raw) + (size_t)GAP * cols;
    return CTP_OK;
}

extern "C" ctp_status ctp_voc_create(ctp_voc** out, const ctp_voc_cfg* c) {
    CTP_REQUIRE(out && c, "voc_create: null argument");
    CTP_REQUIRE(c->dvae_hidden == 256 || c->dvae_hidden == 512, "dvae_hidden must be 256 or 512");
    CTP_REQUIRE(c->voc_dim == 512, "voc_dim must be 512");
    CTP_REQUIRE(c->n_fft == 1024 && c->hop == 256, "ISTFT kernel is built for n_fft 1024 / hop 256");
    CTP_REQUIRE(c->n_mels <= 100 && c->dvae_idim % 8 == 0 && c->dvae_bn % 8 == 0 && c->dvae_odim % 8 == 0, "bad channel counts");
    CTP_REQUIRE(c->max_frames >= 64, "max_frames too small");
    CTP_REQUIRE(c->dvae_dilation * 3 <= GAP, "dilation too large for the inter-utterance gap");
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) { ctp_set_error("no CUDA device: libctp has no CPU fallback"); return CTP_ERR_NO_DEVICE; }
    ctp_status ds = ctp_device_check(dev);
    if (ds != CTP_OK) return ds;
    int gi = gemm_init();
    if (gi) return (ctp_status)gi;
    ctp_voc* h = new ctp_voc();
    h->cfg = *c;
    if (const char* e = getenv("CTP_DWCONV")) h->dw_tile = strcmp(e, "row") != 0;
    if (const char* e = getenv("CTP_GELU")) h->gelu_mode = strcmp(e, "as") == 0 ? 1 : 2;
    const long long R = c->max_frames;
    h->cap_rows = R;
    const int cw = std::max(c->dvae_hidden, c->voc_dim);
    const int iw = std::max(std::max(4 * c->dvae_hidden, c->voc_inter), c->encoder ? c->dvae_idim : 0);
    CTP_REQUIRE(!c->encoder || (c->dvae_odim <= c->n_fft + 2 && c->dvae_odim % 64 == 0), "encoder odim must be a multiple of 64 and <= n_fft + 2");
    int st = 0;
    if (!st) st = voc_alloc(h, &h->x0, R, c->dvae_idim);
    if (!st) st = voc_alloc(h, &h->c1, R, c->dvae_bn);
    if (!st) st = voc_alloc(h, &h->xres, R, cw);
    if (!st) st = voc_alloc(h, &h->y16, R, cw);
    if (!st) st = voc_alloc(h, &h->hm, R, iw);
    if (!st) st = voc_alloc(h, &h->o1, R, c->dvae_odim);
    if (!st) st = voc_alloc(h, &h->mel32, R, c->n_mels);
    if (!st) st = voc_alloc(h, &h->mel16, R, MEL_PAD);
    if (!st) st = voc_alloc(h, &h->head, R, HEAD_PITCH);
    if (!st) st = voc_alloc(h, &h->frames, R, c->n_fft);
    if (!st) st = voc_alloc(h, &h->valid, R, 1);
    if (!st) st = voc_alloc(h, &h->row_src, R, 1);
    if (!st) st = voc_alloc(h, &h->d_row0, h->max_utt, 1);
    if (!st) st = voc_alloc(h, &h->d_nrows, h->max_utt, 1);
    if (!st) st = voc_alloc(h, &h->d_wavoff, h->max_utt, 1);
    if (!st) st = voc_alloc(h, &h->twiddle, 512, 1);
    if (st) { ctp_voc_destroy(h); return (ctp_status)st; }
    std::vector<float2> tw(512);
    for (int k = 0; k < 512; ++k) {
        const double a = 2.0 * M_PI * (double)k / 1024.0;
        tw[k] = make_float2((float)cos(a), (float)sin(a));
    }
    if (cudaMemcpy(h->twiddle, tw.data(), sizeof(float2) * 512, cudaMemcpyHostToDevice) != cudaSuccess) {
        ctp_set_error("voc_create: twiddle upload failed");
        ctp_voc_destroy(h);
        return CTP_ERR_CUDA;
    }
    *out = h;
    return CTP_OK;
}

extern "C" void ctp_voc_destroy(ctp_voc* h) {
    if (!h) return;
    for (void* p : h->allocs) cudaFree(p);
    delete h;
}

extern "C" ctp_status ctp_voc_bind_weights(ctp_voc* h, const ctp_voc_weights* w) {
    CTP_REQUIRE(h && w, "voc_bind: null argument");
    const bool dv = w->conv_in0_w && w->conv_in2_w && w->dvae_blocks && w->conv_out_w && w->out_conv_w && w->coef;
    const bool vc = w->embed_w && w->voc_blocks && w->head_w && w->window && w->norm_w && w->final_ln_w;
    const bool en = h->cfg.encoder && w->conv_in0_w && w->conv_in2_w && w->dvae_blocks && w->conv_out_w && w->coef && w->ds0_w && w->ds0_b &&
                    w->ds2_w && w->ds2_b && w->mel_fb && w->vq_in_w && w->vq_in_b && w->window;
    CTP_REQUIRE(dv || vc || en, "voc_bind: neither a complete DVAE, Vocos nor prompt-encoder weight set was given");
    CTP_REQUIRE(!h->cfg.encoder || en, "voc_bind: encoder handle needs encoder.*, downsample_conv.*, mel filter bank, window, coef and GFSQ project_in");
    CTP_REQUIRE(!dv || !h->cfg.use_vq || (w->vq_proj_w && w->vq_proj_b), "voc_bind: GFSQ projection missing");
    h->w = *w;
    if (dv || en) { h->dvae_blocks.assign(w->dvae_blocks, w->dvae_blocks + h->cfg.dvae_layers); h->w.dvae_blocks = h->dvae_blocks.data(); }
    h->has_encoder = en;
    if (vc) { h->voc_blocks.assign(w->voc_blocks, w->voc_blocks + h->cfg.voc_layers); h->w.voc_blocks = h->voc_blocks.data(); }
    h->has_dvae = dv && !h->cfg.encoder; h->has_vocos = vc;
    h->bound = true;
    return CTP_OK;
}

#define VLAUNCH_OK() do { ctp_count_launch(); cudaError_t _e = cudaGetLastError(); if (_e != cudaSuccess) { ctp_set_error("%s:%d launch: %s", __FILE__, __LINE__, cudaGetErrorString(_e)); return CTP_ERR_CUDA; } } while (0)

// rows x K(window) GEMM against weights [F][K]; window_rows_before = rows the im2col window starts before row r
static int voc_gemm(const __half* act, int act_C, int taps, int rows, const void* wgt, int F, GemmEpilogue e, cudaStream_t s) {
    GemmLaunch g{};
    const int before = taps / 2;
    g.A = act - (long long)before * act_C;
    g.a_rows = rows; g.lda = act_C;
    g.B = wgt; g.b_rows = F; g.ldb = (long long)taps * act_C;
    g.K = (long long)taps * act_C;
    g.block_n = F >= 256 ? 256 : 128;
    g.split_k = 1;
    e.T = rows; e.F = F;
    g.epi = e;
    return gemm_launch(g, s);
}

template <int C>
static int convnext_block(ctp_voc* h, const ctp_convnext_w& b, int rows, int inter, int dil, cudaStream_t s) {
    constexpr int DW_R = 16;
    if (h->dw_tile && (dil == 1 || dil == 2)) k_dwconv_ln_tile<C, DW_R><<<(rows + DW_R - 1) / DW_R, 128, 0, s>>>(h->xres, h->y16, h->valid, b.dw_w, b.dw_b, b.ln_w, b.ln_b, dil, rows);
    else k_dwconv_ln<C><<<rows, 128, 0, s>>>(h->xres, h->y16, h->valid, b.dw_w, b.dw_b, b.ln_w, b.ln_b, dil);
    VLAUNCH_OK();
    GemmEpilogue e1{};
    e1.out = h->hm; e1.ldo = inter; e1.out_f16 = 1; e1.bias = b.pw1_b; e1.act_gelu = h->gelu_mode; e1.row_valid = h->valid;
    int st = voc_gemm(h->y16, C, 1, rows, b.pw1_w, inter, e1, s);
    if (st) return st;
    GemmEpilogue e2{};
    e2.out = h->xres; e2.ldo = C; e2.bias = b.pw2_b; e2.gamma = b.gamma; e2.residual = h->xres; e2.ldr = C; e2.row_valid = h->valid;
    return voc_gemm(h->hm, inter, 1, rows, b.pw2_w, C, e2, s);
}

struct GroupLayout {
    std::vector<int> row0, nrows, row_src;
    std::vector<unsigned char> valid;
    int rows = 0, max_len = 0;
};

// rows of the padded time axis for utterances with `mel_frames[i]` frames; flat0 = index of the first frame of the
// group in the caller's concatenated input
static int voc_layout(ctp_voc* h, int n_utt, const std::vector<int>& mel_frames, long long flat0, const long long* wav_off,
                      GroupLayout& L, cudaStream_t s) {
    const ctp_voc_cfg& c = h->cfg;
    L.row0.resize(n_utt); L.nrows = mel_frames;
    long long R = GAP;
    for (int i = 0; i < n_utt; ++i) { L.row0[i] = (int)R; R += L.nrows[i] + GAP; }
    L.rows = (int)R;
    L.valid.assign(L.rows, 0);
    L.row_src.assign(L.rows, -1);
    long long flat = flat0;
    for (int i = 0; i < n_utt; ++i) {
        for (int f = 0; f < L.nrows[i]; ++f) { L.valid[L.row0[i] + f] = 1; L.row_src[L.row0[i] + f] = (int)(flat + f); }
        flat += L.nrows[i];
        L.max_len = std::max(L.max_len, c.hop * (L.nrows[i] - 1));
    }
    CTP_CUDA_OK(cudaMemcpyAsync(h->valid, L.valid.data(), L.rows, cudaMemcpyHostToDevice, s));
    CTP_CUDA_OK(cudaMemcpyAsync(h->row_src, L.row_src.data(), sizeof(int) * L.rows, cudaMemcpyHostToDevice, s));
    CTP_CUDA_OK(cudaMemcpyAsync(h->d_row0, L.row0.data(), sizeof(int) * n_utt, cudaMemcpyHostToDevice, s));
    CTP_CUDA_OK(cudaMemcpyAsync(h->d_nrows, L.nrows.data(), sizeof(int) * n_utt, cudaMemcpyHostToDevice, s));
    if (wav_off) CTP_CUDA_OK(cudaMemcpyAsync(h->d_wavoff, wav_off, sizeof(long long) * n_utt, cudaMemcpyHostToDevice, s));
    CTP_CUDA_OK(cudaStreamSynchronize(s));  // host vectors are re-used by the caller
    // guard rows after the last row may hold stale data from a longer previous call: conv inputs must see zeros there
    const int rows = L.rows;
    const int cw = std::max(c.dvae_hidden, c.voc_dim);
    CTP_CUDA_OK(cudaMemsetAsync(h->x0 + (long long)rows * c.dvae_idim, 0, sizeof(__half) * GAP * c.dvae_idim, s));
    CTP_CUDA_OK(cudaMemsetAsync(h->c1 + (long long)rows * c.dvae_bn, 0, sizeof(__half) * GAP * c.dvae_bn, s));
    CTP_CUDA_OK(cudaMemsetAsync(h->xres + (long long)rows * cw, 0, sizeof(float) * GAP * cw, s));
    CTP_CUDA_OK(cudaMemsetAsync(h->o1 + (long long)rows * c.dvae_odim, 0, sizeof(__half) * GAP * c.dvae_odim, s));
    CTP_CUDA_OK(cudaMemsetAsync(h->mel16 + (long long)rows * MEL_PAD, 0, sizeof(__half) * GAP * MEL_PAD, s));
    return CTP_OK;
}

static int voc_run_stack(ctp_voc* h, int rows, cudaStream_t s);

// DVAE decode (dvae.py:272-291): src -> mel32 rows
static int voc_run_dvae(ctp_voc* h, const GroupLayout& L, const void* src, cudaStream_t s) {
    const ctp_voc_cfg& c = h->cfg;
    const int rows = L.rows;
    int st;
    if (c.use_vq) {
        k_voc_input_codes<<<rows, 128, 0, s>>>((const int*)src, h->x0, h->row_src, h->w.vq_proj_w, h->w.vq_proj_b, c.dvae_idim, 2, 2);
    } else {
        k_voc_input_hidden<<<rows, 128, 0, s>>>((const float*)src, h->x0, h->row_src, c.dvae_idim);
    }
    VLAUNCH_OK();
    if ((st = voc_run_stack(h, rows, s))) return st;
    {   // conv_out: 1x1, no bias (dvae.py:159,167)
        GemmEpilogue e{};
        e.out = h->o1; e.ldo = c.dvae_odim; e.out_f16 = 1; e.row_valid = h->valid;
        if ((st = voc_gemm(h->y16, c.dvae_hidden, 1, rows, h->w.conv_out_w, c.dvae_odim, e, s))) return st;
    }
    {   // out_conv: Conv1d(dim -> 100, k3, p1, no bias), then * coef (dvae.py:285,291)
        GemmEpilogue e{};
        e.out = h->mel32; e.ldo = c.n_mels; e.gamma = h->w.coef; e.row_valid = h->valid;
        if ((st = voc_gemm(h->o1, c.dvae_odim, 3, rows, h->w.out_conv_w, c.n_mels, e, s))) return st;
    }
    return CTP_OK;
}

// DVAEDecoder.forward up to (not including) conv_out (dvae.py:161-166): x0 rows -> y16 rows (fp16 copy of the residual stream)
static int voc_run_stack(ctp_voc* h, int rows, cudaStream_t s) {
    const ctp_voc_cfg& c = h->cfg;
    int st;
    {   // conv_in[0]: Conv1d(idim -> bn, k3, p1) + GELU   (dvae.py:145-147)
        GemmEpilogue e{};
        e.out = h->c1; e.ldo = c.dvae_bn; e.out_f16 = 1; e.bias = h->w.conv_in0_b; e.act_gelu = h->gelu_mode; e.row_valid = h->valid;
        if ((st = voc_gemm(h->x0, c.dvae_idim, 3, rows, h->w.conv_in0_w, c.dvae_bn, e, s))) return st;
    }
    {   // conv_in[2]: Conv1d(bn -> hidden, k3, p1) -> fp32 residual stream
        GemmEpilogue e{};
        e.out = h->xres; e.ldo = c.dvae_hidden; e.bias = h->w.conv_in2_b; e.row_valid = h->valid;
        if ((st = voc_gemm(h->c1, c.dvae_bn, 3, rows, h->w.conv_in2_w, c.dvae_hidden, e, s))) return st;
    }
    for (int l = 0; l < c.dvae_layers; ++l) {
        if (c.dvae_hidden == 512) st = convnext_block<512>(h, h->dvae_blocks[l], rows, 4 * c.dvae_hidden, c.dvae_dilation, s);
        else st = convnext_block<256>(h, h->dvae_blocks[l], rows, 4 * c.dvae_hidden, c.dvae_dilation, s);
        if (st) return st;
    }
    k_cvt_rows<<<rows, 128, 0, s>>>(h->xres, h->y16, h->valid, c.dvae_hidden, c.dvae_hidden);
    VLAUNCH_OK();
    return CTP_OK;
}

// Vocos decode: mel32 rows -> waveforms
static int voc_run_vocos(ctp_voc* h, const GroupLayout& L, int n_utt, float* wav_out, cudaStream_t s) {
    const ctp_voc_cfg& c = h->cfg;
    const int rows = L.rows;
    int st;
    k_cvt_rows<<<rows, 128, 0, s>>>(h->mel32, h->mel16, h->valid, c.n_mels, MEL_PAD);
    VLAUNCH_OK();
    {   // embed: Conv1d(100 -> 512, k7, p3)
        GemmEpilogue e{};
        e.out = h->xres; e.ldo = c.voc_dim; e.bias = h->w.embed_b; e.row_valid = h->valid;
        if ((st = voc_gemm(h->mel16, MEL_PAD, 7, rows, h->w.embed_w, c.voc_dim, e, s))) return st;
    }
    k_ln_rows<512><<<rows, 128, 0, s>>>(h->xres, h->xres, nullptr, h->valid, h->w.norm_w, h->w.norm_b);
    VLAUNCH_OK();
    for (int l = 0; l < c.voc_layers; ++l)
        if ((st = convnext_block<512>(h, h->voc_blocks[l], rows, c.voc_inter, 1, s))) return st;
    k_ln_rows<512><<<rows, 128, 0, s>>>(h->xres, nullptr, h->y16, h->valid, h->w.final_ln_w, h->w.final_ln_b);
    VLAUNCH_OK();
    {   // ISTFTHead.out: Linear(512 -> n_fft + 2)
        GemmEpilogue e{};
        e.out = h->head; e.ldo = HEAD_PITCH; e.bias = h->w.head_b; e.row_valid = h->valid;
        if ((st = voc_gemm(h->y16, c.voc_dim, 1, rows, h->w.head_w, c.n_fft + 2, e, s))) return st;
    }
    k_istft_frame<<<rows, 256, 0, s>>>(h->head, h->frames, h->valid, h->w.window, h->twiddle);
    VLAUNCH_OK();
    if (L.max_len > 0) {
        k_overlap_add<<<dim3((L.max_len + 255) / 256, n_utt), 256, 0, s>>>(h->frames, wav_out, h->d_row0, h->d_nrows, h->d_wavoff,
                                                                             h->w.window, c.hop, c.n_fft);
        VLAUNCH_OK();
    }
    return CTP_OK;
}

// scatter externally supplied mel rows into the padded layout
__global__ void k_voc_input_mel(const float* __restrict__ mel, float* __restrict__ mel32, const int* __restrict__ row_src, int n_mels) {
    const int r = blockIdx.x;
    const int src = row_src[r];
    for (int c = threadIdx.x; c < n_mels; c += blockDim.x) mel32[(long long)r * n_mels + c] = (src >= 0) ? mel[(long long)src * n_mels + c] : 0.f;
}

// group the utterances so that each group fits the workspace; fn(i, j, first_mel_frame) handles utterances [i, j)
template <typename Fn>
static int voc_for_groups(ctp_voc* h, int n_utt, const std::vector<int>& mel_frames, Fn fn) {
    int i = 0;
    long long frame0 = 0;
    while (i < n_utt) {
        long long R = GAP;
        int j = i;
        while (j < n_utt && j - i < h->max_utt) {
            const long long need = (long long)mel_frames[j] + GAP;
            if (R + need > h->cap_rows) break;
            R += need;
            ++j;
        }
        CTP_REQUIRE(j > i, "voc_decode: utterance %d (%d mel frames) exceeds the workspace (max_frames %d)", i, mel_frames[i], h->cfg.max_frames);
        int st = fn(i, j, frame0);
        if (st) return st;
        for (int k = i; k < j; ++k) frame0 += mel_frames[k];
        i = j;
    }
    return CTP_OK;
}

extern "C" ctp_status ctp_voc_decode(ctp_voc* h, int32_t n_utt, const int32_t* lens_host, const void* src, float* wav_out,
                                     const int64_t* wav_offsets_host, float* mel_out, ctp_stream stream) {
    CTP_REQUIRE(h && h->bound && h->has_dvae, "voc_decode: DVAE weights not bound");
    CTP_REQUIRE(n_utt >= 1 && lens_host && src && (wav_out || mel_out), "voc_decode: bad argument");
    CTP_REQUIRE(!wav_out || (h->has_vocos && wav_offsets_host), "voc_decode: waveform requested but Vocos weights / offsets missing");
    cudaStream_t s = (cudaStream_t)stream;
    std::vector<int> mf(n_utt);
    for (int i = 0; i < n_utt; ++i) { CTP_REQUIRE(lens_host[i] >= 1, "voc_decode: utterance %d has %d frames", i, lens_host[i]); mf[i] = 2 * lens_host[i]; }
    return (ctp_status)voc_for_groups(h, n_utt, mf, [&](int i, int j, long long frame0) -> int {
        std::vector<int> sub(mf.begin() + i, mf.begin() + j);
        std::vector<long long> offs;
        if (wav_out) for (int k = i; k < j; ++k) offs.push_back(wav_offsets_host[k]);
        GroupLayout L;
        int st = voc_layout(h, j - i, sub, frame0, wav_out ? offs.data() : nullptr, L, s);
        if (st) return st;
        if ((st = voc_run_dvae(h, L, src, s))) return st;
        if (mel_out) {
            long long f = frame0;
            for (int k = 0; k < j - i; ++k) {
                CTP_CUDA_OK(cudaMemcpyAsync(mel_out + f * h->cfg.n_mels, h->mel32 + (long long)L.row0[k] * h->cfg.n_mels,
                                            sizeof(float) * (size_t)L.nrows[k] * h->cfg.n_mels, cudaMemcpyDeviceToDevice, s));
                f += L.nrows[k];
            }
        }
        if (wav_out) return voc_run_vocos(h, L, j - i, wav_out, s);
        return CTP_OK;
    });
}

extern "C" ctp_status ctp_voc_encode(ctp_voc* h, int32_t n_samples, const float* audio, int32_t* ids_out, float* feat_out,
                                     int32_t* n_frames_host, ctp_stream stream) {
    CTP_REQUIRE(h && h->bound && h->has_encoder, "voc_encode: no prompt-encoder weights bound (create the handle with encoder = 1)");
    CTP_REQUIRE(audio && ids_out, "voc_encode: null buffer");
    const ctp_voc_cfg& c = h->cfg;
    CTP_REQUIRE(n_samples > c.n_fft / 2, "voc_encode: reflect padding needs more than %d samples (got %d)", c.n_fft / 2, n_samples);
    cudaStream_t s = (cudaStream_t)stream;
    const int T = n_samples / c.hop + 1;          // torch.stft(center=True)
    const int T2 = (T - 2) / 2 + 1;               // Conv1d(k4, stride 2, p1), dvae.py:229
    CTP_REQUIRE(T >= 2, "voc_encode: audio too short");
    CTP_REQUIRE((long long)T + 2 * GAP <= h->cap_rows, "voc_encode: %d mel frames exceed the workspace (%lld rows)", T, h->cap_rows);
    const int dim = c.dvae_idim;
    int st;
    GroupLayout LA;
    if ((st = voc_layout(h, 1, std::vector<int>{T}, 0, nullptr, LA, s))) return (ctp_status)st;
    const int row0 = LA.row0[0];
    // only the T valid rows are written below: the gap rows of this layout must read as zero padding (a longer previous call may have left data)
    CTP_CUDA_OK(cudaMemsetAsync(h->mel16, 0, sizeof(__half) * (size_t)LA.rows * MEL_PAD, s));
    k_mel_frame<<<T, 256, 0, s>>>(audio, n_samples, h->mel16 + (long long)row0 * MEL_PAD, h->w.window, h->twiddle, h->w.mel_fb, h->w.coef,
                                  c.n_mels, c.hop);
    VLAUNCH_OK();
    {   // downsample_conv.0: Conv1d(100 -> dim, k3, p1) + GELU (dvae.py:227-228) -> hm rows (fp16, pitch dim)
        GemmEpilogue e{};
        e.out = h->hm; e.ldo = dim; e.out_f16 = 1; e.bias = h->w.ds0_b; e.act_gelu = h->gelu_mode; e.row_valid = h->valid;
        if ((st = voc_gemm(h->mel16, MEL_PAD, 3, LA.rows, h->w.ds0_w, dim, e, s))) return (ctp_status)st;
    }
    GroupLayout LB;
    if ((st = voc_layout(h, 1, std::vector<int>{T2}, 0, nullptr, LB, s))) return (ctp_status)st;
    CTP_CUDA_OK(cudaMemsetAsync(h->x0, 0, sizeof(__half) * (size_t)LB.rows * dim, s));
    {   // downsample_conv.2: Conv1d(dim -> dim, k4, stride 2, p1) + GELU (dvae.py:229-230): output frame t2 reads input frames
        // 2 t2 - 1 .. 2 t2 + 2, i.e. an im2col window of 4*dim starting one row before row 2 t2 with a row pitch of 2*dim
        GemmLaunch g{};
        g.A = h->hm + (long long)(row0 - 1) * dim; g.a_rows = T2; g.lda = 2LL * dim;
        g.B = h->w.ds2_w; g.b_rows = dim; g.ldb = 4LL * dim; g.K = 4LL * dim;
        g.block_n = dim >= 256 ? 256 : 128; g.split_k = 1;
        g.epi.out = h->x0 + (long long)LB.row0[0] * dim; g.epi.ldo = dim; g.epi.out_f16 = 1; g.epi.bias = h->w.ds2_b; g.epi.act_gelu = h->gelu_mode;
        g.epi.T = T2; g.epi.F = dim;
        if ((st = gemm_launch(g, s))) return (ctp_status)st;
    }
    if ((st = voc_run_stack(h, LB.rows, s))) return (ctp_status)st;
    float* feat = h->head;   // fp32 scratch [rows][>= odim]
    {   // encoder.conv_out: 1x1, no bias (dvae.py:159,167), kept in fp32 for the quantiser
        GemmEpilogue e{};
        e.out = feat; e.ldo = c.dvae_odim; e.row_valid = h->valid;
        if ((st = voc_gemm(h->y16, c.dvae_hidden, 1, LB.rows, h->w.conv_out_w, c.dvae_odim, e, s))) return (ctp_status)st;
    }
    const float* feat0 = feat + (long long)LB.row0[0] * c.dvae_odim;
    k_gfsq_quantize<<<T2, 64, 0, s>>>(feat0, c.dvae_odim, 2, h->w.vq_in_w, h->w.vq_in_b, ids_out);
    VLAUNCH_OK();
    if (feat_out) CTP_CUDA_OK(cudaMemcpyAsync(feat_out, feat0, sizeof(float) * (size_t)T2 * c.dvae_odim, cudaMemcpyDeviceToDevice, s));
    if (n_frames_host) *n_frames_host = T2;
    return CTP_OK;
}

extern "C" ctp_status ctp_voc_quantize(ctp_voc* h, int32_t n_frames, const float* feat, int32_t* ids_out, ctp_stream stream) {
    CTP_REQUIRE(h && h->bound && h->has_encoder, "voc_quantize: no prompt-encoder weights bound (create the handle with encoder = 1)");
    CTP_REQUIRE(n_frames >= 1 && feat && ids_out, "voc_quantize: bad argument");
    k_gfsq_quantize<<<n_frames, 64, 0, (cudaStream_t)stream>>>(feat, h->cfg.dvae_odim, 2, h->w.vq_in_w, h->w.vq_in_b, ids_out);
    VLAUNCH_OK();
    return CTP_OK;
}

extern "C" ctp_status ctp_voc_decode_mel(ctp_voc* h, int32_t n_utt, const int32_t* mel_lens_host, const float* mel, float* wav_out,
                                         const int64_t* wav_offsets_host, ctp_stream stream) {
    CTP_REQUIRE(h && h->bound && h->has_vocos, "voc_decode_mel: Vocos weights not bound");
    CTP_REQUIRE(n_utt >= 1 && mel_lens_host && mel && wav_out && wav_offsets_host, "voc_decode_mel: bad argument");
    cudaStream_t s = (cudaStream_t)stream;
    std::vector<int> mf(mel_lens_host, mel_lens_host + n_utt);
    for (int i = 0; i < n_utt; ++i) CTP_REQUIRE(mf[i] >= 2, "voc_decode_mel: utterance %d has %d mel frames", i, mf[i]);
    return (ctp_status)voc_for_groups(h, n_utt, mf, [&](int i, int j, long long frame0) -> int {
        std::vector<int> sub(mf.begin() + i, mf.begin() + j);
        std::vector<long long> offs;
        for (int k = i; k < j; ++k) offs.push_back(wav_offsets_host[k]);
        GroupLayout L;
        int st = voc_layout(h, j - i, sub, frame0, offs.data(), L, s);
        if (st) return st;
        k_voc_input_mel<<<L.rows, 128, 0, s>>>(mel, h->mel32, h->row_src, h->cfg.n_mels);
        VLAUNCH_OK();
        return voc_run_vocos(h, L, j - i, wav_out, s);
    });
}
