// CLIP image pre-processing on the device: expand2square + bicubic resize + centre crop + rescale + normalise.
//
// What it replaces: t2v_metrics/models/vqascore_models/mm_utils.py:128-139 (expand2square) followed by the CLIP image processor
// of the v3.0 wrapper (resize shortest edge to S with PIL BICUBIC, centre crop SxS, /255, (x - mean) / std), i.e.
// oracle/clipt5_oracle.py:clip_preprocess. The resize is Pillow's (third-party dependency of the reference, not vendored;
// installed 12.2.0, src/libImaging/Resample.c): a separable convolution with per-output-pixel windows, coefficients normalised in
// double and quantised to 22-bit fixed point, the horizontal pass first, its result ROUNDED TO uint8, then the vertical pass. The
// coefficient tables are built on the host in double exactly as Pillow builds them (so the integers are identical), the two integer
// passes run here, and the result is bit-identical to PIL for every pixel.
//
// One CTA = one image x one band of `tile_rows` output rows: the horizontally resampled source rows the band needs live in shared
// memory (uint8), the vertical pass reads them from there, the normalised fp32 / bf16 CHW pixels are written once. The source is
// read once from HBM (plus the band overlap, absorbed by L2), the output written once.
#pragma once
#include <cmath>
#include <cstdint>
#include <string>
#include <vector>
#include "ptx.cuh"

namespace vqa {

constexpr int PRE_PRECISION_BITS = 32 - 8 - 2;      // Resample.c PRECISION_BITS
constexpr int PRE_THREADS = 256;
constexpr int PRE_MAX_SMEM = 200 * 1024;

struct PreImage {            // one per image, lives in the workspace (device) and is built on the host
    long long src_off;       // byte offset of the HWC uint8 RGB image inside `src`
    int h, w;                // stored image
    int paste_x, paste_y;    // where the stored image sits on the (virtual) square canvas; 0 when not padded
    int canvas_h, canvas_w;  // canvas the resize reads (= max(h, w) squared when padded, else h x w)
    int htab, vtab;          // int offsets of the tables inside the workspace: bounds[out][2] then kk[out][ksize]
    int hks, vks;            // taps per output pixel
    int tile_rows;           // output rows per CTA
    int band_rows;           // max source rows any band of this image needs (shared memory rows)
};

// ---------------------------------------------------------------------------------------------- host: Pillow's coefficient tables
inline double pre_bicubic(double x) {                 // Resample.c bicubic_filter, a = -0.5
    const double a = -0.5;
    if (x < 0.0) x = -x;
    if (x < 1.0) return ((a + 2.0) * x - (a + 3.0)) * x * x + 1;
    if (x < 2.0) return (((x - 5) * x + 8) * x - 4) * a;
    return 0.0;
}

// precompute_coeffs + normalize_coeffs_8bpc for output pixels [first, first + count) of an in_size -> out_size resize over the whole
// axis (box = [0, in_size)). Appends bounds (xmin, n) pairs then count * ksize fixed-point taps to `tab`; returns ksize.
inline int pre_build_table(int in_size, int out_size, int first, int count, std::vector<int>& tab) {
    const float in0 = 0.f, in1 = (float)in_size;
    double scale = (double)(in1 - in0) / out_size;
    double filterscale = scale;
    if (filterscale < 1.0) filterscale = 1.0;
    const double support = 2.0 * filterscale;          // BICUBIC.support = 2.0
    const int ksize = (int)std::ceil(support) * 2 + 1;
    const size_t base = tab.size();
    tab.resize(base + (size_t)count * 2 + (size_t)count * ksize, 0);
    int* bounds = tab.data() + base;
    int* kk = bounds + (size_t)count * 2;
    std::vector<double> k(ksize);
    for (int i = 0; i < count; ++i) {
        const int xx = first + i;
        const double center = in0 + (xx + 0.5) * scale;
        double ww = 0.0;
        const double ss = 1.0 / filterscale;
        int xmin = (int)(center - support + 0.5);
        if (xmin < 0) xmin = 0;
        int xmax = (int)(center + support + 0.5);
        if (xmax > in_size) xmax = in_size;
        xmax -= xmin;
        for (int x = 0; x < xmax; ++x) {
            const double w = pre_bicubic((x + xmin - center + 0.5) * ss);
            k[x] = w;
            ww += w;
        }
        for (int x = 0; x < xmax; ++x)
            if (ww != 0.0) k[x] /= ww;
        for (int x = 0; x < ksize; ++x) {
            const double v = x < xmax ? k[x] : 0.0;
            kk[(size_t)i * ksize + x] = v < 0 ? (int)(-0.5 + v * (1 << PRE_PRECISION_BITS)) : (int)(0.5 + v * (1 << PRE_PRECISION_BITS));
        }
        bounds[2 * i] = xmin;
        bounds[2 * i + 1] = xmax;
    }
    return ksize;
}

// ---------------------------------------------------------------------------------------------- device
__device__ __forceinline__ int pre_clip8(int v) {     // Resample.c clip8: table lookup of (v >> PRECISION_BITS) clamped to [0, 255]
    v >>= PRE_PRECISION_BITS;
    return v < 0 ? 0 : (v > 255 ? 255 : v);
}

template <typename OUT>
__global__ void __launch_bounds__(PRE_THREADS)
clip_preprocess_kernel(const uint8_t* __restrict__ src, const PreImage* __restrict__ images, const int* __restrict__ tables,
                       int out_size, uchar3 background, float3 mean, float3 stdv, OUT* __restrict__ out) {
    extern __shared__ uint8_t band[];                 // [band_rows][out_size][3] horizontally resampled rows
    const PreImage im = images[blockIdx.y];
    const int y0 = blockIdx.x * im.tile_rows;
    if (y0 >= out_size) return;
    const int ny = min(im.tile_rows, out_size - y0);
    const int* hb = tables + im.htab;
    const int* hk = hb + 2 * out_size;
    const int* vb = tables + im.vtab;
    const int* vk = vb + 2 * out_size;
    // source rows this band needs: windows are monotonic in y
    const int r0 = vb[2 * y0];
    const int r1 = vb[2 * (y0 + ny - 1)] + vb[2 * (y0 + ny - 1) + 1];
    const int nrows = r1 - r0;
    const int row_elems = out_size * 3;
    const uint8_t* img = src + im.src_off;
    const int bg[3] = {background.x, background.y, background.z};

    // ---- horizontal pass: canvas rows r0 .. r1 -> band (uint8, rounded exactly like ImagingResampleHorizontal_8bpc)
    for (int idx = threadIdx.x; idx < nrows * row_elems; idx += PRE_THREADS) {
        const int r = idx / row_elems, rem = idx - r * row_elems;
        const int xx = rem / 3, c = rem - xx * 3;
        const int xmin = hb[2 * xx], n = hb[2 * xx + 1];
        const int* k = hk + (size_t)xx * im.hks;
        const int sy = r0 + r - im.paste_y;           // row inside the stored image
        int ss = 1 << (PRE_PRECISION_BITS - 1);
        if (sy < 0 || sy >= im.h) {
            for (int x = 0; x < n; ++x) ss += bg[c] * k[x];
        } else {
            const uint8_t* row = img + (size_t)sy * im.w * 3 + c;
            for (int x = 0; x < n; ++x) {
                const int sx = xmin + x - im.paste_x;
                const int v = (sx < 0 || sx >= im.w) ? bg[c] : (int)row[(size_t)sx * 3];
                ss += v * k[x];
            }
        }
        band[idx] = (uint8_t)pre_clip8(ss);
    }
    __syncthreads();

    // ---- vertical pass + rescale + normalise: out[c][y][x] = ((u8 / 255) - mean[c]) / std[c] in fp32 (IEEE division, no fma)
    const float mu[3] = {mean.x, mean.y, mean.z}, sd[3] = {stdv.x, stdv.y, stdv.z};
    OUT* dst = out + (size_t)blockIdx.y * 3 * out_size * out_size;
    for (int idx = threadIdx.x; idx < ny * row_elems; idx += PRE_THREADS) {
        const int xx = idx % out_size;
        const int c = (idx / out_size) % 3;
        const int y = y0 + idx / row_elems;
        const int ymin = vb[2 * y], n = vb[2 * y + 1];
        const int* k = vk + (size_t)y * im.vks;
        const uint8_t* col = band + (size_t)(ymin - r0) * row_elems + xx * 3 + c;
        int ss = 1 << (PRE_PRECISION_BITS - 1);
        for (int j = 0; j < n; ++j) ss += (int)col[(size_t)j * row_elems] * k[j];
        const float u = (float)pre_clip8(ss);
        const float v = __fdiv_rn(__fsub_rn(__fdiv_rn(u, 255.0f), mu[c]), sd[c]);
        const size_t o = ((size_t)c * out_size + y) * out_size + xx;
        if constexpr (sizeof(OUT) == 4) dst[o] = v;
        else dst[o] = __float2bfloat16_rn(v);
    }
}

// ---------------------------------------------------------------------------------------------- host: plan + launch
struct PrePlan {
    std::vector<PreImage> images;
    std::vector<int> tables;
    int max_tiles = 0;
    size_t smem = 0;
    std::string error;
    size_t images_bytes() const { return (images.size() * sizeof(PreImage) + 15) / 16 * 16; }
    size_t bytes() const { return images_bytes() + tables.size() * sizeof(int); }
};

// Geometry of CLIPImageProcessor for one image (shortest edge -> S, int() truncation of the long edge, centre crop S x S), on the
// padded canvas when pad_to_square. Tables are shared between images of the same canvas size.
inline bool pre_plan(const int32_t* heights, const int32_t* widths, const int64_t* offsets, int n, int S, bool pad, PrePlan& plan) {
    struct Key { int ch, cw, htab, vtab, hks, vks, tile, band; };
    std::vector<Key> cache;
    plan.images.resize(n);
    for (int i = 0; i < n; ++i) {
        const int h = heights[i], w = widths[i];
        if (h <= 0 || w <= 0) { plan.error = "image with non-positive size"; return false; }
        PreImage& im = plan.images[i];
        im.src_off = offsets[i]; im.h = h; im.w = w;
        const int side = h > w ? h : w;
        im.canvas_h = pad ? side : h; im.canvas_w = pad ? side : w;
        im.paste_x = pad ? (side - w) / 2 : 0; im.paste_y = pad ? (side - h) / 2 : 0;
        const Key* hit = nullptr;
        for (const Key& k : cache) if (k.ch == im.canvas_h && k.cw == im.canvas_w) { hit = &k; break; }
        if (!hit) {
            const int ch = im.canvas_h, cw = im.canvas_w;
            // get_resize_output_image_size(shortest_edge = S, default_to_square = False): short -> S, long -> int(S * long / short)
            const int nw = cw <= ch ? S : (int)((double)S * cw / ch);
            const int nh = cw <= ch ? (int)((double)S * ch / cw) : S;
            const int left = (nw - S) / 2, top = (nh - S) / 2;        // centre crop (both >= 0 since the short side == S)
            Key k; k.ch = ch; k.cw = cw;
            k.htab = (int)plan.tables.size();
            k.hks = pre_build_table(cw, nw, left, S, plan.tables);
            k.vtab = (int)plan.tables.size();
            k.vks = pre_build_table(ch, nh, top, S, plan.tables);
            // tile rows: as many as fit the shared-memory budget, at most 16
            const int* vb = plan.tables.data() + k.vtab;
            int tile = 16, band = 0;
            for (; tile >= 1; tile >>= 1) {
                band = 0;
                for (int y0 = 0; y0 < S; y0 += tile) {
                    const int y1 = (y0 + tile < S ? y0 + tile : S) - 1;
                    const int rows = vb[2 * y1] + vb[2 * y1 + 1] - vb[2 * y0];
                    if (rows > band) band = rows;
                }
                if ((size_t)band * S * 3 <= (size_t)PRE_MAX_SMEM) break;
            }
            if (tile < 1) { plan.error = "image too large for the device resize (vertical filter window exceeds shared memory)"; return false; }
            k.tile = tile; k.band = band;
            cache.push_back(k);
            hit = &cache.back();
        }
        im.htab = hit->htab; im.vtab = hit->vtab; im.hks = hit->hks; im.vks = hit->vks;
        im.tile_rows = hit->tile; im.band_rows = hit->band;
        const int tiles = (S + im.tile_rows - 1) / im.tile_rows;
        if (tiles > plan.max_tiles) plan.max_tiles = tiles;
        const size_t sm = (size_t)im.band_rows * S * 3;
        if (sm > plan.smem) plan.smem = sm;
    }
    return true;
}

}  // namespace vqa
