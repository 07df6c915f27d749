// TEST INFRASTRUCTURE: the weight builders of speech_editing_toolkit_b200/csrc/mel_frontend_weights.h compiled with g++, so
// that tests/test_mel_frontend_weights.py can emulate the conv-GEMM formulation on the CPU and compare with the oracle.
#include <cstring>

#include "../../speech_editing_toolkit_b200/csrc/mel_frontend_weights.h"

extern "C" {
void mf_dft_weights(int n_fft, int hop, int ndft, float* out) {
  std::vector<float> w;
  fse::melfe::build_dft_weights(n_fft, hop, ndft, w);
  std::memcpy(out, w.data(), w.size() * sizeof(float));
}
void mf_mel_weights(int sr, int n_fft, int n_mels, double fmin, double fmax, int kin, float* out) {
  std::vector<float> w;
  fse::melfe::build_mel_weights(sr, n_fft, n_mels, fmin, fmax, kin, w);
  std::memcpy(out, w.data(), w.size() * sizeof(float));
}
}
