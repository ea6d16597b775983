// B200AsrModel — the reference C++ runtime's `AsrModel` interface (runtime/core/decoder/asr_model.h:20-77) on
// libwenet_b200.so.  Same public methods and semantics as wenet::AsrModel / wenet::TorchAsrModel
// (runtime/core/decoder/torch_asr_model.cc:87-287): chunk-wise ForwardEncoder with the caller-visible feature caching
// of AsrModel::ForwardEncoder / CacheFeature (asr_model.cc:36-66), AttentionRescoring over the encoder output
// accumulated since Reset(), Copy() sharing the (immutable) device weights.  No libtorch, glog or OpenFst: the decoder
// graph search of the reference runtime (ctc_prefix_beam_search.cc, the wfst searcher) keeps calling these two methods.
//
// The model file is written by `python -m wenet_b200.export <train.yaml> <final.pt> <out.wbm>` (wenet_b200/export.py:
// the config struct + the tensors wenet_b200/weights.py packs for wb_model_set_tensor).
#ifndef WENET_B200_RUNTIME_B200_ASR_MODEL_H_
#define WENET_B200_RUNTIME_B200_ASR_MODEL_H_

#include <memory>
#include <string>
#include <vector>

#include "../include/wenet_b200.h"

namespace wenet_b200 {

class B200AsrModel {
 public:
  B200AsrModel() = default;
  B200AsrModel(const B200AsrModel& other);   // shares the device weights, fresh decoding state
  ~B200AsrModel();
  // Loads a .wbm file onto the current CUDA device.  Returns false (and fills error()) on failure.
  bool Read(const std::string& model_path);
  const std::string& error() const { return error_; }

  int right_context() const { return right_context_; }
  int subsampling_rate() const { return subsampling_rate_; }
  int sos() const { return sos_; }
  int eos() const { return eos_; }
  bool is_bidirectional_decoder() const { return is_bidirectional_decoder_; }
  int offset() const { return offset_; }
  int vocab_size() const { return cfg_.vocab; }

  // If chunk_size > 0, streaming case. Otherwise, none streaming case (asr_model.h:29-33)
  void set_chunk_size(int chunk_size) { chunk_size_ = chunk_size; }
  void set_num_left_chunks(int num_left_chunks) { num_left_chunks_ = num_left_chunks; }
  // start: if it is the start chunk of one sentence (asr_model.cc:17-34)
  int num_frames_for_chunk(bool start) const;

  void Reset();
  // chunk_feats: (frames, feature_dim) fbank rows; ctc_prob: (chunk frames, vocab) log-probabilities
  void ForwardEncoder(const std::vector<std::vector<float>>& chunk_feats, std::vector<std::vector<float>>* ctc_prob);
  void AttentionRescoring(const std::vector<std::vector<int>>& hyps, float reverse_weight,
                          std::vector<float>* rescoring_score);
  std::shared_ptr<B200AsrModel> Copy() const;

 private:
  void ForwardEncoderFunc(const std::vector<std::vector<float>>& chunk_feats, std::vector<std::vector<float>>* ctc_prob);
  void CacheFeature(const std::vector<std::vector<float>>& chunk_feats);
  bool Fail(const std::string& what);
  void* Scratch(int which, size_t bytes);   // grow-only device buffers

  std::shared_ptr<wb_model> model_;   // immutable after load; shared between copies
  wb_model_config cfg_ = {};
  int right_context_ = 6;
  int subsampling_rate_ = 4;
  int sos_ = 0, eos_ = 0;
  bool is_bidirectional_decoder_ = false;
  int chunk_size_ = 16;
  int num_left_chunks_ = -1;
  int offset_ = 0;
  std::vector<std::vector<float>> cached_feature_;
  std::string error_;
  // decoding state on the device
  void* stream_ = nullptr;
  float* att_cache_ = nullptr;   // (layers, heads, cache_t1, 128)
  int cache_t1_ = 0;
  size_t att_cache_cap_ = 0;
  float* cnn_cache_ = nullptr;   // (layers, d, kernel - 1); null before the first chunk
  bool have_cnn_cache_ = false;
  float* enc_out_ = nullptr;     // fp32 (frames, d) accumulated since Reset()
  size_t enc_rows_ = 0, enc_cap_ = 0;
  void* scratch_[6] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
  size_t scratch_cap_[6] = {0, 0, 0, 0, 0, 0};
};

}  // namespace wenet_b200

#endif  // WENET_B200_RUNTIME_B200_ASR_MODEL_H_
