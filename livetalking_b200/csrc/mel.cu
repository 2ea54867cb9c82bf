// Wav2Lip mel front-end on the GPU (replaces the CPU librosa/scipy path of
// avatars/wav2lip/audio.py:45-51 as called from MelASR.run_step, avatars/audio_features/mel.py:46-63).
//
//   pre-emphasis (audio.py:20-23)  ->  centre-padded periodic-Hann STFT 800/200 (audio.py:57-61)
//   -> |.| -> 80x401 Slaney mel filterbank (audio.py:92-101) -> 20 log10(max(1e-5, .)) - 20 (audio.py:103-105,47)
//   -> clip(8 (S+100)/100 - 4, -4, 4) (audio.py:110-114) -> B windows of 16 frames (mel.py:50-63)
//
// The reference computes this in float64 (scipy.lfilter promotes); so do we: the whole step is ~54 M double FMAs,
// latency-bound, and float64 keeps the windows within 1e-9 of the CPU path.  The 800-point DFT is evaluated
// directly against a shared-memory twiddle table (800 = 2^5 * 5^2; a radix FFT would save FLOPs nobody is short of).
#include <cmath>
#include <mutex>
#include <vector>

#include "ltb_internal.h"

namespace ltb {

constexpr int kNfft = 800, kHop = 200, kBins = 401, kMels = 80, kMelStep = 16;

// (the three-kernel version kept its spectrum / mel frames in global scratch; the fused kernel needs none)
size_t mel_scratch_spec_doubles(int) { return 2; }
size_t mel_scratch_mel_doubles(int) { return 2; }

// ONE kernel for the whole front-end: one block per STFT frame t.
//   phase 1: pre-emphasis + Hann window of the frame into shared memory, twiddle table
//   phase 2: 401-bin DFT magnitude -> shared memory
//   phase 3: 80 mel bands (thread m), dB, normalise, clip
//   phase 4: the value of (band m, frame t) is written straight into every output window that contains frame t
//            (window i covers frames [start_i, start_i + 16), start_i = int(left + i * mult), tail-clamped: mel.py:50-63)
// Frames that no window reads exit immediately (20 of 84 at B = 16).  No global scratch, one launch instead of three.
__global__ void __launch_bounds__(256) mel_fused_kernel(const float* __restrict__ pcm, int nsamp, const double* __restrict__ fb, int T,
                                                        int B, double left, double mult, float* __restrict__ out) {
  __shared__ double fr[kNfft];
  __shared__ double tc[kNfft];
  __shared__ double ts[kNfft];
  __shared__ double sp[kBins];
  __shared__ int win_lo, win_hi;   // windows [win_lo, win_hi) may contain frame t
  const int t = blockIdx.x;
  auto start_of = [&](int i) {
    int s0 = (int)__dadd_rn(left, __dmul_rn((double)i, mult));
    if (s0 + kMelStep > T) s0 = T - kMelStep;
    return s0;
  };
  if (threadIdx.x == 0) {
    int lo = B, hi = 0;
    for (int i = 0; i < B; ++i) {
      const int s0 = start_of(i);
      if (t >= s0 && t < s0 + kMelStep) {
        lo = min(lo, i);
        hi = max(hi, i + 1);
      }
    }
    win_lo = lo;
    win_hi = hi;
  }
  __syncthreads();
  if (win_lo >= win_hi) return;   // block-uniform: nobody reads this frame
  for (int i = threadIdx.x; i < kNfft; i += 256) {
    const int n = t * kHop + i - kNfft / 2;  // centre padding with zeros
    double y = 0.0;
    if (n >= 0 && n < nsamp) {
      const double x0 = (double)pcm[n];
      const double x1 = n > 0 ? (double)pcm[n - 1] : 0.0;
      y = __dadd_rn(x0, -__dmul_rn(0.97, x1));  // y[n] = x[n] - 0.97 x[n-1]
    }
    const double w = 0.5 - 0.5 * cospi((double)i / 400.0);  // periodic Hann(800)
    fr[i] = __dmul_rn(y, w);
    tc[i] = cospi((double)i / 400.0);
    ts[i] = sinpi((double)i / 400.0);
  }
  __syncthreads();
  for (int k = threadIdx.x; k < kBins; k += 256) {
    double re = 0.0, im = 0.0;
    int idx = 0;
    for (int n = 0; n < kNfft; ++n) {
      re = fma(fr[n], tc[idx], re);
      im = fma(fr[n], ts[idx], im);
      idx += k;
      if (idx >= kNfft) idx -= kNfft;
    }
    sp[k] = sqrt(re * re + im * im);
  }
  __syncthreads();
  const int m = threadIdx.x;
  if (m >= kMels) return;
  const double* f = fb + (size_t)m * kBins;
  double acc = 0.0;
  for (int k = 0; k < kBins; ++k) acc = fma(f[k], sp[k], acc);
  const double min_level = 1e-5;  // exp(-100/20 * ln 10)
  const double db = 20.0 * log10(fmax(min_level, acc)) - 20.0;
  double v = 8.0 * ((db + 100.0) / 100.0) - 4.0;
  v = fmin(4.0, fmax(-4.0, v));
  for (int i = win_lo; i < win_hi; ++i) {
    const int s0 = start_of(i);
    if (t >= s0 && t < s0 + kMelStep) out[((size_t)i * kMels + m) * kMelStep + (t - s0)] = (float)v;
  }
}

// ---- host: Slaney mel filterbank, librosa.filters.mel(sr=16000, n_fft=800, n_mels=80, fmin=55, fmax=7600) ----
static double hz_to_mel(double f) {
  const double f_sp = 200.0 / 3.0, min_log_hz = 1000.0, min_log_mel = min_log_hz / f_sp, logstep = std::log(6.4) / 27.0;
  return f >= min_log_hz ? min_log_mel + std::log(f / min_log_hz) / logstep : f / f_sp;
}
static double mel_to_hz(double m) {
  const double f_sp = 200.0 / 3.0, min_log_hz = 1000.0, min_log_mel = min_log_hz / f_sp, logstep = std::log(6.4) / 27.0;
  return m >= min_log_mel ? min_log_hz * std::exp(logstep * (m - min_log_mel)) : f_sp * m;
}

static double* g_fb_dev[64] = {nullptr};
static std::mutex g_fb_mutex;

static cudaError_t ensure_filterbank(double** out) {
  std::lock_guard<std::mutex> lock(g_fb_mutex);
  int dev = 0;
  cudaError_t e = cudaGetDevice(&dev);
  if (e != cudaSuccess) return e;
  if (dev < 0 || dev >= 64) return cudaErrorInvalidDevice;
  if (!g_fb_dev[dev]) {
    std::vector<double> melf(kMels + 2);
    const double m0 = hz_to_mel(55.0), m1 = hz_to_mel(7600.0);
    for (int i = 0; i < kMels + 2; ++i) melf[i] = mel_to_hz(m0 + (m1 - m0) * i / (kMels + 1));
    std::vector<double> fb((size_t)kMels * kBins);
    for (int i = 0; i < kMels; ++i) {
      const double enorm = 2.0 / (melf[i + 2] - melf[i]);
      for (int k = 0; k < kBins; ++k) {
        const double f = 8000.0 * k / (kBins - 1);
        const double lower = (f - melf[i]) / (melf[i + 1] - melf[i]);
        const double upper = (melf[i + 2] - f) / (melf[i + 2] - melf[i + 1]);
        double w = std::fmin(lower, upper);
        if (w < 0) w = 0;
        fb[(size_t)i * kBins + k] = (double)(float)(w * enorm);  // librosa returns float32
      }
    }
    double* d = nullptr;
    e = cudaMalloc(&d, fb.size() * sizeof(double));
    if (e != cudaSuccess) return e;
    e = cudaMemcpy(d, fb.data(), fb.size() * sizeof(double), cudaMemcpyHostToDevice);
    if (e != cudaSuccess) return e;
    g_fb_dev[dev] = d;
  }
  *out = g_fb_dev[dev];
  return cudaSuccess;
}

cudaError_t launch_mel_step(const float* pcm, int nsamp, int B, int stride_left_chunks, int fps, double* scratch_spec,
                            double* scratch_mel, float* out, cudaStream_t st) {
  double* fb = nullptr;
  cudaError_t e = ensure_filterbank(&fb);
  if (e != cudaSuccess) return e;
  const int T = 1 + nsamp / kHop;
  if (T < kMelStep) return cudaErrorInvalidValue;
  (void)scratch_spec;   // the fused kernel keeps the spectrum / mel frame in shared memory
  (void)scratch_mel;
  const double left = (double)(stride_left_chunks * 80) / 50.0;  // mel.py:50
  const double mult = 80.0 / (double)fps;                        // mel.py:52
  mel_fused_kernel<<<T, 256, 0, st>>>(pcm, nsamp, fb, T, B, left, mult, out);
  return cudaGetLastError();
}

}  // namespace ltb
