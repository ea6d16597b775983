// See b200_asr_model.h.  Host orchestration only: every compute step is a libwenet_b200.so entry point.
#include "b200_asr_model.h"

#include <cuda_runtime.h>

#include <algorithm>
#include <cstdio>
#include <cstring>
#include <limits>

namespace wenet_b200 {

namespace {
constexpr char kMagic[8] = {'W', 'B', 'M', '0', '0', '0', '1', '\0'};

struct ModelDeleter {
  void operator()(wb_model* m) const { wb_model_destroy(m); }
};

inline size_t RoundUp(size_t x, size_t a) { return (x + a - 1) / a * a; }
}  // namespace

bool B200AsrModel::Fail(const std::string& what) {
  error_ = what + ": " + (wb_last_error() ? wb_last_error() : "");
  return false;
}

void* B200AsrModel::Scratch(int which, size_t bytes) {
  if (scratch_cap_[which] < bytes) {
    if (scratch_[which]) cudaFree(scratch_[which]);
    scratch_[which] = nullptr;
    const size_t cap = RoundUp(bytes + bytes / 4 + 1024, 256);
    if (cudaMalloc(&scratch_[which], cap) != cudaSuccess) {
      scratch_cap_[which] = 0;
      return nullptr;
    }
    scratch_cap_[which] = cap;
  }
  return scratch_[which];
}

bool B200AsrModel::Read(const std::string& model_path) {
  FILE* f = fopen(model_path.c_str(), "rb");
  if (!f) {
    error_ = "cannot open " + model_path;
    return false;
  }
  char magic[8];
  int32_t hdr[4];   // sos, eos, bidirectional, tensor count
  bool ok = fread(magic, 1, 8, f) == 8 && memcmp(magic, kMagic, 8) == 0 && fread(&cfg_, sizeof(cfg_), 1, f) == 1 &&
            fread(hdr, sizeof(int32_t), 4, f) == 4;
  if (!ok) {
    fclose(f);
    error_ = "bad model file header";
    return false;
  }
  wb_model* raw = nullptr;
  if (wb_model_create(&raw, &cfg_) != WB_OK) {
    fclose(f);
    return Fail("wb_model_create");
  }
  model_ = std::shared_ptr<wb_model>(raw, ModelDeleter());
  std::vector<char> buf;
  for (int i = 0; i < hdr[3] && ok; ++i) {
    int32_t name_len = 0, dtype = 0;
    int64_t numel = 0;
    ok = fread(&name_len, 4, 1, f) == 1 && name_len > 0 && name_len < 256;
    std::string name(ok ? name_len : 0, '\0');
    ok = ok && fread(&name[0], 1, name_len, f) == (size_t)name_len && fread(&dtype, 4, 1, f) == 1 &&
         fread(&numel, 8, 1, f) == 1 && numel > 0;
    if (!ok) break;
    const size_t bytes = (size_t)numel * (dtype == WB_BF16 ? 2 : 4);
    buf.resize(bytes);
    ok = fread(buf.data(), 1, bytes, f) == bytes;
    if (ok && wb_model_set_tensor(raw, name.c_str(), buf.data(), dtype, numel) != WB_OK) {
      fclose(f);
      return Fail("wb_model_set_tensor(" + name + ")");
    }
  }
  fclose(f);
  if (!ok) {
    error_ = "truncated model file";
    return false;
  }
  cudaStream_t st;
  if (cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking) != cudaSuccess) {
    error_ = "cudaStreamCreate failed";
    return false;
  }
  stream_ = st;
  if (wb_model_finalize(raw, stream_) != WB_OK) return Fail("wb_model_finalize");
  sos_ = hdr[0];
  eos_ = hdr[1];
  is_bidirectional_decoder_ = hdr[2] != 0 && cfg_.rdec_layers > 0;
  Reset();
  return true;
}

B200AsrModel::B200AsrModel(const B200AsrModel& other)
    : model_(other.model_), cfg_(other.cfg_), right_context_(other.right_context_),
      subsampling_rate_(other.subsampling_rate_), sos_(other.sos_), eos_(other.eos_),
      is_bidirectional_decoder_(other.is_bidirectional_decoder_), chunk_size_(other.chunk_size_),
      num_left_chunks_(other.num_left_chunks_) {
  // inner states for forward are not copied (torch_asr_model.cc:87-104); every copy decodes on its own stream
  cudaStream_t st = nullptr;
  cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking);
  stream_ = st;
}

B200AsrModel::~B200AsrModel() {
  if (stream_) cudaStreamSynchronize((cudaStream_t)stream_);
  cudaFree(att_cache_);
  cudaFree(cnn_cache_);
  cudaFree(enc_out_);
  for (void* p : scratch_) cudaFree(p);
  if (stream_) cudaStreamDestroy((cudaStream_t)stream_);
}

std::shared_ptr<B200AsrModel> B200AsrModel::Copy() const {
  auto m = std::make_shared<B200AsrModel>(*this);
  m->Reset();   // reset the inner states for new decoding (torch_asr_model.cc:106-111)
  return m;
}

int B200AsrModel::num_frames_for_chunk(bool start) const {
  int num_required_frames = 0;
  if (chunk_size_ > 0) {
    if (!start) {   // first batch
      const int context = right_context_ + 1;   // add current frame
      num_required_frames = (chunk_size_ - 1) * subsampling_rate_ + context;
    } else {
      num_required_frames = chunk_size_ * subsampling_rate_;
    }
  } else {
    num_required_frames = std::numeric_limits<int>::max();
  }
  return num_required_frames;
}

void B200AsrModel::Reset() {
  offset_ = 0;
  cache_t1_ = 0;
  have_cnn_cache_ = false;
  enc_rows_ = 0;
  cached_feature_.clear();
}

void B200AsrModel::CacheFeature(const std::vector<std::vector<float>>& chunk_feats) {
  const int cached_feature_size = 1 + right_context_ - subsampling_rate_;
  if ((int)chunk_feats.size() >= cached_feature_size) {
    cached_feature_.resize(cached_feature_size);
    for (int i = 0; i < cached_feature_size; ++i)
      cached_feature_[i] = chunk_feats[chunk_feats.size() - cached_feature_size + i];
  }
}

void B200AsrModel::ForwardEncoder(const std::vector<std::vector<float>>& chunk_feats,
                                  std::vector<std::vector<float>>* ctc_prob) {
  ctc_prob->clear();
  const int num_frames = (int)(cached_feature_.size() + chunk_feats.size());
  if (num_frames >= right_context_ + 1) {
    ForwardEncoderFunc(chunk_feats, ctc_prob);
    CacheFeature(chunk_feats);
  }
}

void B200AsrModel::ForwardEncoderFunc(const std::vector<std::vector<float>>& chunk_feats,
                                      std::vector<std::vector<float>>* out_prob) {
  cudaStream_t st = (cudaStream_t)stream_;
  const int d = cfg_.d_model, V = cfg_.vocab, L = cfg_.enc_layers, H = cfg_.heads;
  const int lead = cfg_.cnn_causal ? cfg_.cnn_kernel - 1 : 0;
  // 1. splice cached_feature_ and chunk_feats (torch_asr_model.cc:121-143)
  const int T = (int)(cached_feature_.size() + chunk_feats.size());
  const int fdim = (int)chunk_feats[0].size();
  std::vector<float> host((size_t)T * fdim);
  for (size_t i = 0; i < cached_feature_.size(); ++i) memcpy(&host[i * fdim], cached_feature_[i].data(), fdim * sizeof(float));
  for (size_t i = 0; i < chunk_feats.size(); ++i)
    memcpy(&host[(cached_feature_.size() + i) * fdim], chunk_feats[i].data(), fdim * sizeof(float));
  float* xs = (float*)Scratch(0, host.size() * sizeof(float));
  cudaMemcpyAsync(xs, host.data(), host.size() * sizeof(float), cudaMemcpyHostToDevice, st);
  // 2. encoder chunk forward (forward_encoder_chunk, asr_model.py:385-426)
  const int required_cache_size = chunk_size_ * num_left_chunks_;
  const int chunk = T >= 7 ? ((T - 1) / 2 - 1) / 2 : 0;
  if (chunk <= 0) return;
  const int key_size = cache_t1_ + chunk;
  int nxt = 0;
  if (required_cache_size == 0) nxt = key_size;
  else if (required_cache_size > 0) nxt = std::max(key_size - required_cache_size, 0);
  const int new_t1 = key_size - nxt;
  float* y = (float*)Scratch(1, (size_t)chunk * d * sizeof(float));
  float* r_att = (float*)Scratch(2, (size_t)L * H * std::max(new_t1, 1) * 128 * sizeof(float));
  float* r_cnn = (float*)Scratch(3, (size_t)L * d * std::max(lead, 1) * sizeof(float));
  const size_t wsb = wb_encoder_chunk_workspace_bytes(model_.get(), T, cache_t1_);
  void* ws = Scratch(4, wsb);
  int oc = 0, on = 0;
  if (wb_encoder_forward_chunk(model_.get(), xs, T, offset_, required_cache_size, cache_t1_ > 0 ? att_cache_ : nullptr,
                               cache_t1_, have_cnn_cache_ ? cnn_cache_ : nullptr, y, r_att, lead > 0 ? r_cnn : nullptr, &oc,
                               &on, ws, wsb, stream_) != WB_OK) {
    Fail("wb_encoder_forward_chunk");
    return;
  }
  // the returned caches become the next call's inputs (swap the buffers)
  std::swap(att_cache_, *reinterpret_cast<float**>(&scratch_[2]));
  std::swap(att_cache_cap_, scratch_cap_[2]);
  cache_t1_ = on;
  if (lead > 0) {
    if (!cnn_cache_) cudaMalloc((void**)&cnn_cache_, (size_t)L * d * lead * sizeof(float));
    cudaMemcpyAsync(cnn_cache_, r_cnn, (size_t)L * d * lead * sizeof(float), cudaMemcpyDeviceToDevice, st);
    have_cnn_cache_ = true;
  }
  offset_ += oc;
  // keep the encoder output for AttentionRescoring (encoder_outs_, torch_asr_model.cc:176-183)
  if (enc_rows_ + oc > enc_cap_) {
    const size_t cap = std::max<size_t>(2 * enc_cap_, enc_rows_ + oc + 1024);
    float* nb = nullptr;
    cudaMalloc((void**)&nb, cap * d * sizeof(float));
    if (enc_rows_) cudaMemcpyAsync(nb, enc_out_, enc_rows_ * d * sizeof(float), cudaMemcpyDeviceToDevice, st);
    cudaStreamSynchronize(st);
    cudaFree(enc_out_);
    enc_out_ = nb;
    enc_cap_ = cap;
  }
  cudaMemcpyAsync(enc_out_ + enc_rows_ * d, y, (size_t)oc * d * sizeof(float), cudaMemcpyDeviceToDevice, st);
  enc_rows_ += oc;
  // 3. ctc_activation (asr_model.py:428-438)
  const int p3 = cfg_.precise ? 3 : 1;
  const int64_t ldl = (V + 7) / 8 * 8;
  char* s5 = (char*)Scratch(5, RoundUp((size_t)oc * d * p3 * 2, 256) + (size_t)oc * ldl * 4 + (size_t)oc * 8 + 256);
  void* a16 = s5;
  float* logp = (float*)(s5 + RoundUp((size_t)oc * d * p3 * 2, 256));
  float* tv = logp + (size_t)oc * ldl;
  int32_t* ti = (int32_t*)(tv + oc);
  if (wb_op_cast_bf16(y, d, oc, d, a16, (int64_t)d * p3, cfg_.precise, stream_) != WB_OK ||
      wb_ctc_logprobs(model_.get(), a16, oc, 0, 0.f, logp, ldl, 1, tv, ti, stream_) != WB_OK) {
    Fail("ctc_activation");
    return;
  }
  std::vector<float> hp((size_t)oc * ldl);
  cudaMemcpyAsync(hp.data(), logp, hp.size() * sizeof(float), cudaMemcpyDeviceToHost, st);
  cudaStreamSynchronize(st);
  out_prob->resize(oc);
  for (int i = 0; i < oc; ++i) (*out_prob)[i].assign(hp.begin() + (size_t)i * ldl, hp.begin() + (size_t)i * ldl + V);
}

void B200AsrModel::AttentionRescoring(const std::vector<std::vector<int>>& hyps, float reverse_weight,
                                      std::vector<float>* rescoring_score) {
  const int num_hyps = (int)hyps.size();
  rescoring_score->assign(num_hyps, 0.0f);
  if (num_hyps == 0 || enc_rows_ == 0 || cfg_.dec_layers == 0) return;   // no hypothesis / no encoder output
  cudaStream_t st = (cudaStream_t)stream_;
  const int d = cfg_.d_model, V = cfg_.vocab;
  const int p3 = cfg_.precise ? 3 : 1;
  const int64_t ldl = (V + 7) / 8 * 8;
  std::vector<int32_t> hyp_utt(num_hyps, 0), hyp_len(num_hyps), hyp_tok0(num_hyps), toks;
  int64_t R = 0;
  for (int i = 0; i < num_hyps; ++i) {
    hyp_len[i] = (int32_t)hyps[i].size();
    hyp_tok0[i] = (int32_t)toks.size();
    toks.insert(toks.end(), hyps[i].begin(), hyps[i].end());
    R += hyp_len[i] + 1;
  }
  if (toks.empty()) toks.push_back(0);
  const bool use_r2l = is_bidirectional_decoder_ && reverse_weight > 0;
  const int32_t seq_start = 0, seq_len = (int32_t)enc_rows_;
  const size_t enc16 = RoundUp(enc_rows_ * d * p3 * 2, 256), lp_bytes = RoundUp((size_t)R * ldl * 4, 256);
  const size_t wsb = wb_rescoring_workspace_bytes(model_.get(), (int64_t)enc_rows_, R);
  char* s = (char*)Scratch(5, enc16 + 2 * lp_bytes + wsb + 1024);
  void* a16 = s;
  float* lp = (float*)(s + enc16);
  float* rlp = (float*)(s + enc16 + lp_bytes);
  void* ws = s + enc16 + 2 * lp_bytes;
  if (wb_op_cast_bf16(enc_out_, d, (int)enc_rows_, d, a16, (int64_t)d * p3, cfg_.precise, stream_) != WB_OK ||
      wb_decoder_logprobs(model_.get(), a16, (int64_t)enc_rows_, &seq_start, &seq_len, 1, num_hyps, hyp_utt.data(),
                          hyp_len.data(), hyp_tok0.data(), toks.data(), sos_, eos_, use_r2l ? 1 : 0, lp,
                          use_r2l ? rlp : nullptr, ldl, ws, wsb, stream_) != WB_OK) {
    Fail("forward_attention_decoder");
    return;
  }
  std::vector<float> h((size_t)R * ldl), rh;
  cudaMemcpyAsync(h.data(), lp, h.size() * sizeof(float), cudaMemcpyDeviceToHost, st);
  if (use_r2l) {
    rh.resize((size_t)R * ldl);
    cudaMemcpyAsync(rh.data(), rlp, rh.size() * sizeof(float), cudaMemcpyDeviceToHost, st);
  }
  cudaStreamSynchronize(st);
  // ComputeAttentionScore (torch_asr_model.cc:196-206) for both directions, then the combination of :283-285
  int64_t row0 = 0;
  for (int i = 0; i < num_hyps; ++i) {
    const std::vector<int>& hyp = hyps[i];
    const int n = (int)hyp.size();
    float score = 0.0f, r_score = 0.0f;
    for (int j = 0; j < n; ++j) score += h[(size_t)(row0 + j) * ldl + hyp[j]];
    score += h[(size_t)(row0 + n) * ldl + eos_];
    if (use_r2l) {
      for (int j = 0; j < n; ++j) r_score += rh[(size_t)(row0 + j) * ldl + hyp[n - 1 - j]];
      r_score += rh[(size_t)(row0 + n) * ldl + eos_];
    }
    (*rescoring_score)[i] = score * (1 - reverse_weight) + r_score * reverse_weight;
    row0 += n + 1;
  }
}

}  // namespace wenet_b200
