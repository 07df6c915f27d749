// Host-side construction of the two weight matrices of the mel front-end (mel_frontend.cu), plain C++ so that
// tests/tools/mel_frontend_host.cpp can build the very same matrices with g++ and tests/test_mel_frontend_weights.py can check
// the conv-GEMM formulation (taps, centring, zero padding, Slaney filterbank) against the oracle on the CPU-only container.
#pragma once
#include <cmath>
#include <vector>

namespace fse {
namespace melfe {

inline double hz_to_mel(double f) {       // Slaney scale (librosa htk=False)
  const double f_sp = 200.0 / 3, min_log_hz = 1000.0, min_log_mel = min_log_hz / f_sp, logstep = std::log(6.4) / 27.0;
  return f >= min_log_hz ? min_log_mel + std::log(f / min_log_hz) / logstep : f / f_sp;
}
inline double mel_to_hz(double m) {
  const double f_sp = 200.0 / 3, min_log_hz = 1000.0, min_log_mel = min_log_hz / f_sp, logstep = std::log(6.4) / 27.0;
  return m >= min_log_mel ? min_log_hz * std::exp(logstep * (m - min_log_mel)) : f_sp * m;
}

// Windowed DFT as conv weights w[n][c][r] (n < ndft columns, c < hop samples of a row, r < R = n_fft / hop taps):
// column 2k = window * cos, 2k + 1 = -window * sin of bin k at sample s = r * hop + c of the frame; tap r reads row t + r - R/2.
inline void build_dft_weights(int n_fft, int hop, int ndft, std::vector<float>& w) {
  const int R = n_fft / hop, nbins = n_fft / 2 + 1;
  const double two_pi = 6.283185307179586476925286766559;
  w.assign(static_cast<size_t>(ndft) * hop * R, 0.f);
  for (int k = 0; k < nbins; ++k)
    for (int r = 0; r < R; ++r)
      for (int c = 0; c < hop; ++c) {
        const int s = r * hop + c;
        const double win = 0.5 - 0.5 * std::cos(two_pi * s / n_fft);                         // periodic Hann
        const double ang = two_pi * (static_cast<long long>(k) * s % n_fft) / n_fft;
        w[(static_cast<size_t>(2 * k) * hop + c) * R + r] = static_cast<float>(win * std::cos(ang));
        w[(static_cast<size_t>(2 * k + 1) * hop + c) * R + r] = static_cast<float>(-win * std::sin(ang));
      }
}

// librosa.filters.mel(sr, n_fft, n_mels, fmin, fmax), htk=False, norm='slaney', as w[i][k] over a magnitude row of kin >= n_fft/2+1 entries
inline void build_mel_weights(int sample_rate, int n_fft, int n_mels, double fmin, double fmax, int kin, std::vector<float>& w) {
  const int nbins = n_fft / 2 + 1;
  std::vector<double> edges(n_mels + 2);
  const double m0 = hz_to_mel(fmin), m1 = hz_to_mel(fmax);
  for (int i = 0; i < n_mels + 2; ++i) edges[i] = mel_to_hz(m0 + (m1 - m0) * i / (n_mels + 1));
  w.assign(static_cast<size_t>(n_mels) * kin, 0.f);
  for (int i = 0; i < n_mels; ++i) {
    const double enorm = 2.0 / (edges[i + 2] - edges[i]);
    for (int k = 0; k < nbins; ++k) {
      const double f = (sample_rate / 2.0) * k / (nbins - 1);
      const double lower = (f - edges[i]) / (edges[i + 1] - edges[i]), upper = (edges[i + 2] - f) / (edges[i + 2] - edges[i + 1]);
      w[static_cast<size_t>(i) * kin + k] = static_cast<float>(std::fmax(0.0, std::fmin(lower, upper)) * enorm);
    }
  }
}

}  // namespace melfe
}  // namespace fse
