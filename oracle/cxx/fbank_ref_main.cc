// TEST INFRASTRUCTURE ONLY (oracle/_ref/fbank_ref): a command-line driver around the REFERENCE's own C++ front-end,
// compiled from the sources where they lie (/root/reference/runtime/core/frontend/{fbank.h,fft.h,fft.cc}; nothing of the
// reference is copied into this repository).  It exposes what the reference's `wenet::Fbank` computes so that
// tests/test_oracle_pin.py can pin oracle/wenet_oracle.py against it:
//
//   fbank_ref melscale  <slaney|htk> f0 f1 ...                 -> "freq mel inverse(mel)" per line (fbank.h:176-218)
//   fbank_ref filters   <slaney|htk> <num_bins> <sample_rate> <frame_length> <low_freq>
//                                                              -> per bin: "bin first_index n w0 w1 ..." (fbank.h:91-150)
//   fbank_ref fbank     <kaldi|whisper> <num_bins> <pcm.f32>   -> num_frames lines of num_bins values (fbank.h:247-326);
//                       kaldi   = the runtime's fbank configuration (feature_pipeline.h:55-63: povey, HTK mel, ln, low_freq 20,
//                                 pre-emphasis, input in int16 range)
//                       whisper = feature_pipeline.h:64-73 (hanning, slaney, log10, Whisper normalisation, input scaled to [-1, 1))
//
// `bins_` is a private member of wenet::Fbank; the driver reads it through the usual test-only access trick.
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <sstream>   // (before the access trick: the standard library must see its own headers untouched)
#include <stdexcept>
#include <string>
#include <vector>
#include <random>
#include <limits>
#include <utility>
#include <cmath>

#define private public
#include "frontend/fbank.h"
#undef private

static wenet::MelType mel_type_of(const char* s) {
  if (!std::strcmp(s, "slaney")) return wenet::MelType::kSlaney;
  if (!std::strcmp(s, "htk")) return wenet::MelType::kHTK;
  std::fprintf(stderr, "unknown mel type %s\n", s);
  std::exit(2);
}

int main(int argc, char** argv) {
  if (argc < 3) {
    std::fprintf(stderr, "usage: see the header comment of oracle/cxx/fbank_ref_main.cc\n");
    return 2;
  }
  const std::string cmd = argv[1];
  if (cmd == "melscale") {
    const wenet::MelType mt = mel_type_of(argv[2]);
    for (int i = 3; i < argc; ++i) {
      const float f = std::strtof(argv[i], nullptr);
      const float m = wenet::Fbank::MelScale(f, mt);
      std::printf("%.9g %.9g %.9g\n", f, m, wenet::Fbank::InverseMelScale(m, mt));
    }
    return 0;
  }
  if (cmd == "filters") {
    if (argc != 7) return 2;
    const wenet::MelType mt = mel_type_of(argv[2]);
    const int num_bins = std::atoi(argv[3]), sr = std::atoi(argv[4]), flen = std::atoi(argv[5]);
    const float low = std::strtof(argv[6], nullptr);
    wenet::Fbank fb(num_bins, sr, flen, flen / 2 > 0 ? flen / 2 : 1, low, false, false, 1e-10f, wenet::LogBase::kBase10,
                    wenet::WindowType::kHanning, mt, wenet::NormalizationType::kKaldi);
    for (int b = 0; b < num_bins; ++b) {
      std::printf("%d %d %zu", b, fb.bins_[b].first, fb.bins_[b].second.size());
      for (float w : fb.bins_[b].second) std::printf(" %.9g", w);
      std::printf("\n");
    }
    return 0;
  }
  if (cmd == "fbank") {
    if (argc != 5) return 2;
    const bool whisper = !std::strcmp(argv[2], "whisper");
    const int num_bins = std::atoi(argv[3]);
    FILE* f = std::fopen(argv[4], "rb");
    if (!f) {
      std::perror(argv[4]);
      return 1;
    }
    std::vector<float> wave;
    float buf[4096];
    size_t n;
    while ((n = std::fread(buf, sizeof(float), 4096, f)) > 0) wave.insert(wave.end(), buf, buf + n);
    std::fclose(f);
    // the two configurations of runtime/core/frontend/feature_pipeline.h:50-80 (frame 25 ms / shift 10 ms at 16 kHz)
    wenet::Fbank fb = whisper ? wenet::Fbank(num_bins, 16000, 400, 160, 0.0f, false, true, 1e-10f, wenet::LogBase::kBase10,
                                             wenet::WindowType::kHanning, wenet::MelType::kSlaney,
                                             wenet::NormalizationType::kWhisper)
                              : wenet::Fbank(num_bins, 16000, 400, 160);
    std::vector<std::vector<float>> feat;
    const int frames = fb.Compute(wave, &feat);
    for (int i = 0; i < frames; ++i) {
      for (int j = 0; j < num_bins; ++j) std::printf(j ? " %.9g" : "%.9g", feat[i][j]);
      std::printf("\n");
    }
    return 0;
  }
  std::fprintf(stderr, "unknown command %s\n", cmd.c_str());
  return 2;
}
