// Batch assembly on the device: the reference keeps a chunk of pixels on the host, lets a DataLoader pick
// `batch_size` of them per step through ImageChunk.__getitem__ (data/PreSight/my_dataset.py:52-73) and copies the batch to the
// GPU (my_datamanager.py:257-285).  With 180 GB of HBM the chunk itself lives on the device and one kernel gathers the step's
// rows of every field and forms the ray indices (image, pixel // width, pixel % width).
#include "common.cuh"

namespace ps {

struct BatchFields {
    const float* rgbs;            // [n,3]
    const uint8_t* segs;          // [n]
    const float* skies;           // [n]
    const float* depths;          // [n]
    const float* features;        // [n,C] nullable
    const int64_t* pixel_indices; // [n]
    const int64_t* image_indices; // [n]
    const int64_t* video_ids;     // [n]
    const int64_t* widths;        // [n]
    int C;
    // outputs, [B, ...] in batch order
    float* rgb;
    uint8_t* seg;
    float* sky;
    float* depth;
    float* feat;
    int64_t* image_index;
    int64_t* video_id;
    int64_t* ray_index;           // [B,3]
};

// one warp per batch row: lanes stride over the feature channels (coalesced 128-byte rows), lane 0 moves the scalars
__global__ void __launch_bounds__(256) assemble_batch_kernel(BatchFields f, const int64_t* __restrict__ idx, int64_t B,
                                                             int64_t n_chunk, int* __restrict__ bad) {
    const int lane = threadIdx.x & 31;
    const int64_t b = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (b >= B) return;
    const int64_t i = idx[b];
    if (i < 0 || i >= n_chunk) {            // the reference's indexing would raise: report; the row gets a harmless ray index
        if (lane == 0) {
            atomicExch(bad, 1);
            f.ray_index[b * 3] = 0; f.ray_index[b * 3 + 1] = 0; f.ray_index[b * 3 + 2] = 0;
            f.image_index[b] = 0; f.video_id[b] = 0;
        }
        return;
    }
    if (f.features) {
        const float* src = f.features + i * f.C;
        float* dst = f.feat + b * f.C;
        for (int c = lane; c < f.C; c += 32) dst[c] = __ldg(src + c);
    }
    if (lane < 3) f.rgb[b * 3 + lane] = __ldg(f.rgbs + i * 3 + lane);
    if (lane == 3) {
        if (f.seg) f.seg[b] = f.segs[i];
        f.sky[b] = __ldg(f.skies + i);
        f.depth[b] = __ldg(f.depths + i);
    }
    if (lane == 4) {
        const int64_t img = f.image_indices[i], pix = f.pixel_indices[i], w = f.widths[i];
        f.image_index[b] = img;
        f.video_id[b] = f.video_ids[i];
        f.ray_index[b * 3] = img;
        f.ray_index[b * 3 + 1] = pix / w;
        f.ray_index[b * 3 + 2] = pix % w;
    }
}

}  // namespace ps

using namespace ps;

extern "C" int ps_assemble_batch(const float* rgbs, const uint8_t* segs, const float* skies, const float* depths,
                                 const float* features, int C, const int64_t* pixel_indices, const int64_t* image_indices,
                                 const int64_t* video_ids, const int64_t* widths, int64_t n_chunk, const int64_t* idx,
                                 int64_t B, float* rgb, uint8_t* seg, float* sky, float* depth, float* feat,
                                 int64_t* image_index, int64_t* video_id, int64_t* ray_index, int* bad_index_flag,
                                 void* stream) {
    if (B == 0) return 0;
    PS_REQUIRE(rgbs && skies && depths && pixel_indices && image_indices && video_ids && widths && idx,
               "assemble_batch: null input");
    PS_REQUIRE(rgb && sky && depth && image_index && video_id && ray_index && bad_index_flag, "assemble_batch: null output");
    PS_REQUIRE((features == nullptr) == (feat == nullptr) && (segs != nullptr || seg == nullptr),
               "assemble_batch: features / seg inputs and outputs must come in pairs");
    PS_REQUIRE(features == nullptr || C >= 1, "assemble_batch: %d feature channels", C);
    PS_REQUIRE(n_chunk >= 1 && B >= 0, "assemble_batch: empty chunk");
    BatchFields f{rgbs, segs, skies, depths, features, pixel_indices, image_indices, video_ids, widths, C,
                  rgb, seg, sky, depth, feat, image_index, video_id, ray_index};
    const int64_t blocks = cdiv(B * 32, 256);
    PS_REQUIRE(blocks < (1ll << 31), "assemble_batch: batch too large");
    assemble_batch_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(f, idx, B, n_chunk, bad_index_flag);
    return check_launch("assemble_batch");
}
