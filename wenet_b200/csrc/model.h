// Internal model representation behind the opaque wb_model handle.
#pragma once
#include "common.cuh"
#include "kernels.h"
#include "../../include/wenet_b200.h"
#include <map>
#include <string>
#include <vector>

namespace wb {

struct DevTensor {
    void* ptr = nullptr;
    int dtype = 0;
    int64_t numel = 0;
};

struct Linear {
    const void* w = nullptr;   // bf16 [N][K]
    const float* b = nullptr;  // [N] or null
    int N = 0, K = 0;
    WeightMaps tmap;
};

struct Norm {
    const float* g = nullptr;
    const float* b = nullptr;
};

struct EncLayer {
    Norm n_ffm, n_mha, n_conv, n_ff, n_final;
    Linear ffm1, ffm2, ff1, ff2, qkv, out, pw1, pw2;
    const float* pos_u = nullptr;
    const float* pos_v = nullptr;
    const void* pos_w3 = nullptr;  // bf16 [d][3d] packed hi|hi|lo
    float* pos_proj = nullptr;     // [max_pos][d] fp32, built by finalize
    const float* dw_w = nullptr;
    const float* dw_b = nullptr;
    Norm n_cnn;                    // LayerNorm gamma/beta or folded BatchNorm scale/shift
    const float* pad_vec = nullptr;
};

struct DecLayer {
    Norm n1, n2, n3;
    Linear sa_qkv, sa_out, ca_q, ca_kv, ca_out, ff1, ff2;
};

struct Decoder {
    const float* emb = nullptr;  // [V][d]
    std::vector<DecLayer> layers;
    Norm after;
    Linear out;
};

struct Model {
    wb_model_config cfg;
    std::map<std::string, DevTensor> tensors;
    bool finalized = false;
    int F1 = 0, F2 = 0;
    // encoder front
    const float* cmvn_mean = nullptr;
    const float* cmvn_istd = nullptr;
    const float* conv1_w = nullptr;
    const float* conv1_b = nullptr;
    Linear conv2, embed_out;
    const float* pe = nullptr;  // [max_pos][d]
    std::vector<EncLayer> layers;
    Norm after;
    Linear ctc;
    Decoder left, right;
    std::vector<void*> owned;  // extra device allocations made by finalize
};

int model_get(const Model* m, const std::string& name, int dtype, int64_t numel, const void** out);

}  // namespace wb
