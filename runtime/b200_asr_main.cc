// Command-line driver of B200AsrModel for the tests: streams a feature matrix through ForwardEncoder exactly the way the
// reference decoder does (runtime/core/decoder/asr_decoder.cc AdvanceDecoding: num_frames_for_chunk(start) frames per
// call), then rescoring.  Usage:
//   b200_asr_main model.wbm feats.bin out.bin chunk_size num_left_chunks reverse_weight hyps.txt
// feats.bin : int32 frames, int32 dim, then frames*dim float32.   hyps.txt : one hypothesis per line, space separated ids.
// out.bin   : int32 rows, int32 vocab, rows*vocab float32 CTC log-probs, int32 n_hyps, n_hyps float32 rescoring scores.
#include <cstdio>
#include <cstdlib>
#include <fstream>
#include <sstream>
#include <string>
#include <vector>

#include "b200_asr_model.h"

int main(int argc, char** argv) {
  if (argc < 8) {
    fprintf(stderr, "usage: %s model.wbm feats.bin out.bin chunk_size num_left_chunks reverse_weight hyps.txt\n", argv[0]);
    return 2;
  }
  wenet_b200::B200AsrModel base;
  if (!base.Read(argv[1])) {
    fprintf(stderr, "load failed: %s\n", base.error().c_str());
    return 1;
  }
  base.set_chunk_size(atoi(argv[4]));
  base.set_num_left_chunks(atoi(argv[5]));
  auto model = base.Copy();   // decode on a copy, as the reference's decoder threads do
  FILE* f = fopen(argv[2], "rb");
  int32_t frames = 0, dim = 0;
  if (!f || fread(&frames, 4, 1, f) != 1 || fread(&dim, 4, 1, f) != 1) return 1;
  std::vector<std::vector<float>> feats(frames, std::vector<float>(dim));
  for (auto& r : feats)
    if (fread(r.data(), 4, dim, f) != (size_t)dim) return 1;
  fclose(f);
  std::vector<std::vector<float>> all_prob;
  int pos = 0;
  bool start = false;
  while (pos < frames) {
    const int want = model->num_frames_for_chunk(start);
    const int n = std::min(want, frames - pos);
    std::vector<std::vector<float>> chunk(feats.begin() + pos, feats.begin() + pos + n), prob;
    model->ForwardEncoder(chunk, &prob);
    if (!model->error().empty()) {
      fprintf(stderr, "forward failed: %s\n", model->error().c_str());
      return 1;
    }
    all_prob.insert(all_prob.end(), prob.begin(), prob.end());
    pos += n;
    start = true;
  }
  std::vector<std::vector<int>> hyps;
  std::ifstream hf(argv[7]);
  std::string line;
  while (std::getline(hf, line)) {
    std::istringstream is(line);
    std::vector<int> h;
    int t;
    while (is >> t) h.push_back(t);
    hyps.push_back(h);
  }
  std::vector<float> scores;
  model->AttentionRescoring(hyps, (float)atof(argv[6]), &scores);
  FILE* o = fopen(argv[3], "wb");
  const int32_t rows = (int32_t)all_prob.size(), V = model->vocab_size(), nh = (int32_t)scores.size();
  fwrite(&rows, 4, 1, o);
  fwrite(&V, 4, 1, o);
  for (auto& r : all_prob) fwrite(r.data(), 4, V, o);
  fwrite(&nh, 4, 1, o);
  fwrite(scores.data(), 4, nh, o);
  fclose(o);
  printf("frames %d -> %d ctc rows, offset %d, %d hyps rescored\n", frames, rows, model->offset(), nh);
  return 0;
}
