#pragma once
#include "net.h"

namespace sdod {

class VaeDecoder : public NetBase {
public:
    VaeDecoder(const WeightStore* ws, unsigned long long seed, int latent_hw, int max_batch);
    ~VaeDecoder() override;
    int decode(cudaStream_t s, const float* z, uint8_t* image_u8, float* image_f32, int B, bool use_graph);
    int latent_hw() const { return hw_; }

private:
    Act res_block(const Act& x, const std::string& prefix, int cout);
    Act attn_block(const Act& x, const std::string& prefix);
    std::unique_ptr<Plan> build(int B);

    int hw_, max_batch_;
    float* z_in_ = nullptr;       // [maxB, hw, hw, 4] fp32
    float* conv_out_ = nullptr;   // [maxB, 8hw, 8hw, 3] fp32 (pre-clamp)
    uint8_t* u8_out_ = nullptr;
    float* img_out_ = nullptr;
    std::map<int, std::unique_ptr<Plan>> plans_;
};

}  // namespace sdod
