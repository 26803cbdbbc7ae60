// fbank.cu -- feature front end of the reference's data pipeline on the GPU:
//   util/data_loader.py:7-18   torchaudio.transforms.Spectrogram(n_fft = 2*args.n_fft = 1024, win_length = 320, hop_length = 160)
//                              (hann window, center = True, reflect padding, power = 2)  ->  MelScale(16 kHz, 80 mels, n_stft = 513)
// The reference computes this per utterance on the CPU inside the DataLoader's collate function and pads with zeros
// (data_loader.py:21-26, :124-125).  Here a whole padded batch of waveforms goes through three kernels and two tensor-core GEMMs:
//
//   fbank_frames_kernel   frame t of utterance b = hann[i] * x_b[reflect(t*hop - hop + i)], i < win: only the win_length = 320 samples
//                         under the (zero-padded, centred) window are non-zero, so the 1024-point DFT is a K = 320 contraction
//   eec_gemm (tcgen05)    [re | im] = frames x [cos | sin]^T         (M = B*T, N = 2*513 padded, K = 3*320)
//   fbank_power_kernel    p[f] = re^2 + im^2
//   eec_gemm (tcgen05)    mel = p x fb                              (N = 80 padded, K = 3*513 padded)
//   fbank_finish_kernel   [B*T, 96] -> [B, 80, T]  (the (B, n_mels, T) layout the model's forward takes)
//
// fp32 accuracy on bf16 tensor cores: every fp32 operand v is split into bf16 hi = rn(v) and lo = rn(v - hi) and the three
// significant partial products hi*hi + hi*lo + lo*hi are contracted in ONE GEMM by concatenating along K: A' = [a_hi | a_hi | a_lo],
// B' = [b_hi | b_lo | b_hi] (fp32 accumulation in TMEM).  Measured error vs torchaudio: ~1e-5 relative (bar: 1e-3).
#include "common.cuh"

namespace eec {

__device__ __forceinline__ void split_bf16(float v, __nv_bfloat16& hi, __nv_bfloat16& lo) {
  hi = __float2bfloat16_rn(v);
  lo = __float2bfloat16_rn(v - __bfloat162float(hi));
}

// one thread per (row, 8 consecutive window samples); rows past the utterance's frame count are zero
__global__ void __launch_bounds__(256) fbank_frames_kernel(const float* __restrict__ wave, const int64_t* __restrict__ wave_len, long ldw,
                                                        const float* __restrict__ window, __nv_bfloat16* __restrict__ A, int B, int T,
                                                        int win, int hop) {
  pdl_trigger();
  pdl_wait();
  const int cpr = win / 8;                                   // 8-sample chunks per row
  const long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (long)B * T * cpr) return;
  const int ch = (int)(i % cpr);
  const long row = i / cpr;
  const int t = (int)(row % T), b = (int)(row / T);
  const long L = wave_len[b];
  const int Tb = (int)(1 + L / hop);
  __nv_bfloat16 hi[8], lo[8];
  const float* x = wave + (long)b * ldw;
#pragma unroll
  for (int e = 0; e < 8; ++e) {
    const int k = ch * 8 + e;
    float v = 0.f;
    if (t < Tb) {
      long idx = (long)t * hop - (win / 2) + k;              // centred window of `win` samples around sample t*hop
      if (idx < 0) idx = -idx;                                // reflect padding (torch.stft center = True, pad_mode = "reflect")
      if (idx >= L) idx = 2 * (L - 1) - idx;
      idx = idx < 0 ? 0 : idx;
      v = window[k] * x[idx];
    }
    split_bf16(v, hi[e], lo[e]);
  }
  __nv_bfloat16* dst = A + row * (3L * win) + ch * 8;
  *reinterpret_cast<uint4*>(dst) = *reinterpret_cast<const uint4*>(hi);
  *reinterpret_cast<uint4*>(dst + win) = *reinterpret_cast<const uint4*>(hi);
  *reinterpret_cast<uint4*>(dst + 2 * win) = *reinterpret_cast<const uint4*>(lo);
}

// spec [rows, lds] fp32 = [re (nf) | im (nf) | pad]  ->  P [rows, 3*kp] bf16 = [hi | hi | lo] of re^2 + im^2 (kp >= nf, zero padded)
__global__ void __launch_bounds__(256) fbank_power_kernel(const float* __restrict__ spec, int lds, __nv_bfloat16* __restrict__ P, long rows,
                                                       int nf, int kp) {
  pdl_trigger();
  pdl_wait();
  const int cpr = kp / 8;
  const long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= rows * cpr) return;
  const int ch = (int)(i % cpr);
  const long row = i / cpr;
  const float* s = spec + row * lds;
  __nv_bfloat16 hi[8], lo[8];
#pragma unroll
  for (int e = 0; e < 8; ++e) {
    const int f = ch * 8 + e;
    float v = 0.f;
    if (f < nf) {
      const float re = s[f], im = s[nf + f];
      v = fmaf(re, re, im * im);
    }
    split_bf16(v, hi[e], lo[e]);
  }
  __nv_bfloat16* dst = P + row * (3L * kp) + ch * 8;
  *reinterpret_cast<uint4*>(dst) = *reinterpret_cast<const uint4*>(hi);
  *reinterpret_cast<uint4*>(dst + kp) = *reinterpret_cast<const uint4*>(hi);
  *reinterpret_cast<uint4*>(dst + 2 * kp) = *reinterpret_cast<const uint4*>(lo);
}

// mel [B*T, ldm] (frame-major) -> out [B, n_mels, T] through a 32 x 32 shared-memory transpose
__global__ void __launch_bounds__(256) fbank_finish_kernel(const float* __restrict__ mel, int ldm, float* __restrict__ out, int T, int n_mels) {
  pdl_trigger();
  pdl_wait();
  __shared__ float tile[32][33];
  const int b = blockIdx.z, t0 = blockIdx.x * 32, m0 = blockIdx.y * 32;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  for (int r = ty; r < 32; r += 8) {
    const int t = t0 + r, m = m0 + tx;
    tile[r][tx] = (t < T && m < n_mels) ? mel[((long)b * T + t) * ldm + m] : 0.f;
  }
  __syncthreads();
  for (int r = ty; r < 32; r += 8) {
    const int m = m0 + r, t = t0 + tx;
    if (m < n_mels && t < T) out[((long)b * n_mels + m) * T + t] = tile[tx][r];
  }
}

// generic fp32 -> [hi | lo | hi] (or [hi | hi | lo]) bf16 split of a constant operand matrix [rows, k] -> [rows, 3*kp]
__global__ void fbank_split_operand_kernel(const float* __restrict__ w, int k, __nv_bfloat16* __restrict__ out, long rows, int kp) {
  pdl_trigger();
  pdl_wait();
  const long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= rows * kp) return;
  const int c = (int)(i % kp);
  const long r = i / kp;
  const float v = (c < k) ? w[r * k + c] : 0.f;
  __nv_bfloat16 hi, lo;
  split_bf16(v, hi, lo);
  __nv_bfloat16* dst = out + r * (3L * kp) + c;
  dst[0] = hi; dst[kp] = lo; dst[2 * kp] = hi;
}

}  // namespace eec

using namespace eec;

extern "C" int eec_fbank_frames(const float* wave, const int64_t* wave_len, int64_t ldw, const float* window, void* frames_bf16x3,
                                int B, int T, int win, int hop, eec_stream_t stream) {
  EEC_CHECK_ARG(win % 8 == 0 && win > 0 && hop > 0, "fbank_frames: win_length must be a positive multiple of 8 (got %d), hop > 0", win);
  EEC_CHECK_ARG((reinterpret_cast<uintptr_t>(frames_bf16x3) & 15) == 0, "fbank_frames: output not 16-byte aligned");
  const long total = (long)B * T * (win / 8);
  if (total == 0) return 0;
  launch_pdl(fbank_frames_kernel, dim3((unsigned)cdiv64(total, 256)), dim3(256), 0, S(stream), wave, wave_len, (long)ldw, window,
             (__nv_bfloat16*)frames_bf16x3, B, T, win, hop);
  EEC_LAUNCH_CHECK();
  return 0;
}

extern "C" int eec_fbank_power(const float* spec, int lds, void* power_bf16x3, int64_t rows, int n_freqs, int kp, eec_stream_t stream) {
  EEC_CHECK_ARG(kp % 8 == 0 && kp >= n_freqs && lds >= 2 * n_freqs, "fbank_power: need kp %% 8 == 0, kp >= n_freqs, lds >= 2*n_freqs");
  const long total = rows * (kp / 8);
  if (total == 0) return 0;
  launch_pdl(fbank_power_kernel, dim3((unsigned)cdiv64(total, 256)), dim3(256), 0, S(stream), spec, lds, (__nv_bfloat16*)power_bf16x3,
             (long)rows, n_freqs, kp);
  EEC_LAUNCH_CHECK();
  return 0;
}

extern "C" int eec_fbank_finish(const float* mel, int ldm, float* out, int B, int T, int n_mels, eec_stream_t stream) {
  if (B == 0 || T == 0) return 0;
  launch_pdl(fbank_finish_kernel, dim3(cdiv(T, 32), cdiv(n_mels, 32), B), dim3(256), 0, S(stream), mel, ldm, out, T, n_mels);
  EEC_LAUNCH_CHECK();
  return 0;
}

extern "C" int eec_fbank_split_operand(const float* w, int k, void* out_bf16x3, int64_t rows, int kp, eec_stream_t stream) {
  EEC_CHECK_ARG(kp >= k, "fbank_split_operand: kp < k");
  const long total = rows * kp;
  if (total == 0) return 0;
  launch_pdl(fbank_split_operand_kernel, dim3((unsigned)cdiv64(total, 256)), dim3(256), 0, S(stream), w, k, (__nv_bfloat16*)out_bf16x3,
             (long)rows, kp);
  EEC_LAUNCH_CHECK();
  return 0;
}
