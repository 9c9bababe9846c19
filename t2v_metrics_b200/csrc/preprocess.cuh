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
// One CTA = one image x one band of `tile_rows` output rows. The canvas rows the band needs are staged `chunk_rows` at a time into
// shared memory as packed RGBX words (coalesced byte reads, padding colour filled in, so the passes need no bounds checks), the
// horizontal pass turns them into RGBX words of the band buffer (one coefficient load serves up to 4 rows x 3 channels), the vertical
// pass reads the band (one word = 3 channels) and a 768-entry lookup table does the /255, -mean, /std in exact fp32. The source is read
// once from HBM (plus the band overlap, absorbed by L2), the output written once.
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
    int out_h, out_w;        // output window (after the crop) of this image
    long long out_off;       // element offset of this image's output inside `out`
    int x_lo, span;          // canvas columns [x_lo, x_lo + span) the horizontal windows of this image touch
    int chunk_rows;          // canvas rows staged per pass (1..8)
};

enum PreLayout { PRE_CHW = 0, PRE_QWEN_PATCHES = 1 };
struct PrePatchGeom { int patch, merge, temporal; };     // PRE_QWEN_PATCHES only

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

// LAYOUT PRE_CHW: out[c][y][x] per image ([3, out_h, out_w]). PRE_QWEN_PATCHES: the Qwen2-VL processor's patch rows
// (image_processing_qwen2_vl.py:191-220): row ((by * gw/m + bx) * m + iy) * m + ix for the 14x14 patch at grid (by*m+iy, bx*m+ix),
// column (c * temporal + t) * ps*ps + py * ps + px, the still frame written to every temporal slot t.
constexpr int PRE_XL = 128;                 // x lanes; PRE_THREADS / PRE_XL row groups
constexpr int PRE_RG = PRE_THREADS / PRE_XL;
constexpr int PRE_MAX_CHUNK = 4 * PRE_RG;   // rows per staging pass: each thread keeps up to 4 rows x 3 channels of accumulators

template <typename OUT, int LAYOUT>
__global__ void __launch_bounds__(PRE_THREADS)
image_preprocess_kernel(const uint8_t* __restrict__ src, const PreImage* __restrict__ images, const int* __restrict__ tables,
                        uchar3 background, float3 mean, float3 stdv, PrePatchGeom geom, OUT* __restrict__ out) {
    extern __shared__ uint32_t pre_smem[];
    const PreImage im = images[blockIdx.y];
    const int y0 = blockIdx.x * im.tile_rows;
    if (y0 >= im.out_h) return;
    const int out_w = im.out_w;
    const int ny = min(im.tile_rows, im.out_h - y0);
    float* lut = reinterpret_cast<float*>(pre_smem);                 // [3][256] normalised value of every grey level
    uint32_t* band = pre_smem + 768;                                  // [band_rows][out_w] RGBX, horizontally resampled canvas rows
    uint32_t* stage = band + (size_t)im.band_rows * out_w;            // [chunk_rows][span] RGBX canvas rows
    const int* hb = tables + im.htab;
    const int* hk = hb + 2 * out_w;
    const int* vb = tables + im.vtab;
    const int* vk = vb + 2 * im.out_h;
    // canvas rows this band needs: the windows are monotonic in y
    const int r0 = vb[2 * y0];
    const int r1 = vb[2 * (y0 + ny - 1)] + vb[2 * (y0 + ny - 1) + 1];
    const int nrows = r1 - r0;
    const uint8_t* img = src + im.src_off;
    const uint32_t bgw = (uint32_t)background.x | ((uint32_t)background.y << 8) | ((uint32_t)background.z << 16);
    const int lane_x = threadIdx.x % PRE_XL, grp = threadIdx.x / PRE_XL;

    // out = ((u8 / 255) - mean[c]) / std[c] in fp32 with IEEE division and no fma, as numpy / torch compute it on the host
    for (int i = threadIdx.x; i < 768; i += PRE_THREADS) {
        const int c = i >> 8;
        const float mu = c == 0 ? mean.x : (c == 1 ? mean.y : mean.z), sd = c == 0 ? stdv.x : (c == 1 ? stdv.y : stdv.z);
        lut[i] = __fdiv_rn(__fsub_rn(__fdiv_rn((float)(i & 255), 255.0f), mu), sd);
    }

    for (int rc = 0; rc < nrows; rc += im.chunk_rows) {
        const int nr = min(im.chunk_rows, nrows - rc);
        // ---- stage canvas rows r0+rc .. as RGBX words; outside the pasted image the canvas is the padding colour
        for (int rr = 0; rr < nr; ++rr) {
            const int sy = r0 + rc + rr - im.paste_y;
            const bool row_in = sy >= 0 && sy < im.h;
            const uint8_t* row = img + (size_t)(row_in ? sy : 0) * im.w * 3;
            for (int cx = threadIdx.x; cx < im.span; cx += PRE_THREADS) {
                const int sx = cx + im.x_lo - im.paste_x;
                uint32_t px = bgw;
                if (row_in && sx >= 0 && sx < im.w) {
                    const uint8_t* p = row + (size_t)sx * 3;
                    px = (uint32_t)p[0] | ((uint32_t)p[1] << 8) | ((uint32_t)p[2] << 16);
                }
                stage[rr * im.span + cx] = px;
            }
        }
        __syncthreads();
        // ---- horizontal pass (ImagingResampleHorizontal_8bpc): thread = one output column, rows grp, grp + PRE_RG, ...
        for (int xx = lane_x; xx < out_w; xx += PRE_XL) {
            const int xmin = hb[2 * xx] - im.x_lo, n = hb[2 * xx + 1];
            const int* k = hk + (size_t)xx * im.hks;
            int acc[4][3];
#pragma unroll
            for (int q = 0; q < 4; ++q) acc[q][0] = acc[q][1] = acc[q][2] = 1 << (PRE_PRECISION_BITS - 1);
            for (int x = 0; x < n; ++x) {
                const int kv = __ldg(k + x);
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    const int rr = grp + PRE_RG * q;
                    if (rr < nr) {
                        const uint32_t px = stage[rr * im.span + xmin + x];
                        acc[q][0] += (int)(px & 0xffu) * kv;
                        acc[q][1] += (int)((px >> 8) & 0xffu) * kv;
                        acc[q][2] += (int)((px >> 16) & 0xffu) * kv;
                    }
                }
            }
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                const int rr = grp + PRE_RG * q;
                if (rr < nr)
                    band[(size_t)(rc + rr) * out_w + xx] = (uint32_t)pre_clip8(acc[q][0]) | ((uint32_t)pre_clip8(acc[q][1]) << 8) |
                                                           ((uint32_t)pre_clip8(acc[q][2]) << 16);
            }
        }
        __syncthreads();
    }

    // ---- vertical pass (ImagingResampleVertical_8bpc) + normalisation
    OUT* dst = out + im.out_off;
    const int ps = geom.patch, mg = geom.merge, pp = geom.patch * geom.patch;
    const int gwm = LAYOUT == PRE_QWEN_PATCHES ? out_w / (ps * mg) : 0;
    for (int yy = grp; yy < ny; yy += PRE_RG) {
        const int y = y0 + yy;
        const int ymin = vb[2 * y] - r0, n = vb[2 * y + 1];
        const int* k = vk + (size_t)y * im.vks;
        for (int xx = lane_x; xx < out_w; xx += PRE_XL) {
            int a0 = 1 << (PRE_PRECISION_BITS - 1), a1 = a0, a2 = a0;
            const uint32_t* col = band + (size_t)ymin * out_w + xx;
            for (int j = 0; j < n; ++j) {
                const int kv = __ldg(k + j);
                const uint32_t px = col[(size_t)j * out_w];
                a0 += (int)(px & 0xffu) * kv;
                a1 += (int)((px >> 8) & 0xffu) * kv;
                a2 += (int)((px >> 16) & 0xffu) * kv;
            }
            const float v[3] = {lut[pre_clip8(a0)], lut[256 + pre_clip8(a1)], lut[512 + pre_clip8(a2)]};
            if constexpr (LAYOUT == PRE_CHW) {
#pragma unroll
                for (int c = 0; c < 3; ++c) {
                    const size_t o = ((size_t)c * im.out_h + y) * out_w + xx;
                    if constexpr (sizeof(OUT) == 4) dst[o] = v[c];
                    else dst[o] = __float2bfloat16_rn(v[c]);
                }
            } else {
                const int gy = y / ps, py = y - gy * ps, gx = xx / ps, px_ = xx - gx * ps;
                const size_t row = (((size_t)(gy / mg) * gwm + gx / mg) * mg + gy % mg) * mg + gx % mg;
                OUT* q = dst + row * (size_t)(3 * geom.temporal * pp) + py * ps + px_;
#pragma unroll
                for (int c = 0; c < 3; ++c)
                    for (int t = 0; t < geom.temporal; ++t) {
                        if constexpr (sizeof(OUT) == 4) q[(size_t)(c * geom.temporal + t) * pp] = v[c];
                        else q[(size_t)(c * geom.temporal + t) * pp] = __float2bfloat16_rn(v[c]);
                    }
            }
        }
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

struct PreTableKey { int ch, cw, nh, nw, top, left, oh, ow, htab, vtab, hks, vks, tile, band, x_lo, span, chunk; };
inline size_t pre_smem_bytes(int band_rows, int out_w, int chunk_rows, int span) {
    return 768 * sizeof(float) + ((size_t)band_rows * out_w + (size_t)chunk_rows * span) * sizeof(uint32_t);
}

// Tables + band geometry for "resize the ch x cw canvas to nh x nw, keep the oh x ow window at (top, left)"; shared between images
// with the same geometry. Returns nullptr (plan.error set) when the vertical window cannot fit the shared-memory budget.
inline const PreTableKey* pre_tables(PrePlan& plan, std::vector<PreTableKey>& cache, int ch, int cw, int nh, int nw, int top, int left,
                                     int oh, int ow) {
    for (const PreTableKey& k : cache)
        if (k.ch == ch && k.cw == cw && k.nh == nh && k.nw == nw && k.top == top && k.left == left && k.oh == oh && k.ow == ow) return &k;
    PreTableKey k{ch, cw, nh, nw, top, left, oh, ow, 0, 0, 0, 0, 0, 0, 0, 0, 0};
    k.htab = (int)plan.tables.size();
    k.hks = pre_build_table(cw, nw, left, ow, plan.tables);
    k.vtab = (int)plan.tables.size();
    k.vks = pre_build_table(ch, nh, top, oh, plan.tables);
    const int* hb = plan.tables.data() + k.htab;
    const int* vb = plan.tables.data() + k.vtab;
    k.x_lo = hb[0];                                // horizontal windows are monotonic: first start .. last end
    k.span = hb[2 * (ow - 1)] + hb[2 * (ow - 1) + 1] - k.x_lo;
    // output rows per CTA (<= 16) and canvas rows staged per pass (<= PRE_MAX_CHUNK): the largest that fit the shared-memory budget
    int tile = 16, band = 0, chunk = 1;
    for (; tile >= 1; tile >>= 1) {
        band = 0;
        for (int y0 = 0; y0 < oh; y0 += tile) {
            const int y1 = (y0 + tile < oh ? y0 + tile : oh) - 1;
            const int rows = vb[2 * y1] + vb[2 * y1 + 1] - vb[2 * y0];
            if (rows > band) band = rows;
        }
        for (chunk = PRE_MAX_CHUNK; chunk > 1 && pre_smem_bytes(band, ow, chunk, k.span) > (size_t)PRE_MAX_SMEM; chunk >>= 1) {}
        if (pre_smem_bytes(band, ow, chunk, k.span) <= (size_t)PRE_MAX_SMEM) break;
    }
    if (tile < 1) { plan.error = "image too large for the device resize (filter windows exceed shared memory)"; return nullptr; }
    k.tile = tile; k.band = band; k.chunk = chunk;
    cache.push_back(k);
    return &cache.back();
}

inline void pre_finish_image(PrePlan& plan, PreImage& im, const PreTableKey& k) {
    im.htab = k.htab; im.vtab = k.vtab; im.hks = k.hks; im.vks = k.vks;
    im.tile_rows = k.tile; im.band_rows = k.band; im.out_h = k.oh; im.out_w = k.ow;
    im.x_lo = k.x_lo; im.span = k.span; im.chunk_rows = k.chunk;
    const int tiles = (k.oh + k.tile - 1) / k.tile;
    if (tiles > plan.max_tiles) plan.max_tiles = tiles;
    const size_t sm = pre_smem_bytes(k.band, k.ow, k.chunk, k.span);
    if (sm > plan.smem) plan.smem = sm;
}

// Geometry of CLIPImageProcessor for one image (shortest edge -> S, int() truncation of the long edge, centre crop S x S), on the
// padded canvas when pad_to_square. Output [n, 3, S, S].
inline bool pre_plan(const int32_t* heights, const int32_t* widths, const int64_t* offsets, int n, int S, bool pad, PrePlan& plan) {
    std::vector<PreTableKey> cache;
    cache.reserve(n);
    plan.images.resize(n);
    for (int i = 0; i < n; ++i) {
        const int h = heights[i], w = widths[i];
        if (h <= 0 || w <= 0) { plan.error = "image with non-positive size"; return false; }
        PreImage& im = plan.images[i];
        im.src_off = offsets[i]; im.h = h; im.w = w;
        const int side = h > w ? h : w;
        im.canvas_h = pad ? side : h; im.canvas_w = pad ? side : w;
        im.paste_x = pad ? (side - w) / 2 : 0; im.paste_y = pad ? (side - h) / 2 : 0;
        const int ch = im.canvas_h, cw = im.canvas_w;
        // get_resize_output_image_size(shortest_edge = S, default_to_square = False): short -> S, long -> int(S * long / short)
        const int nw = cw <= ch ? S : (int)((double)S * cw / ch);
        const int nh = cw <= ch ? (int)((double)S * ch / cw) : S;
        const PreTableKey* k = pre_tables(plan, cache, ch, cw, nh, nw, (nh - S) / 2, (nw - S) / 2, S, S);   // centre crop
        if (!k) return false;
        pre_finish_image(plan, im, *k);
        im.out_off = (long long)i * 3 * S * S;
    }
    return true;
}

// qwen_vl_utils / Qwen2VLImageProcessor smart_resize (image_processing_qwen2_vl.py:62-87): both sides multiples of `factor`, pixel
// count within [min_pixels, max_pixels]. Python's round() is round-half-even = nearbyint in the default rounding mode.
inline bool pre_smart_resize(int height, int width, int factor, long long min_pixels, long long max_pixels, int& h_bar, int& w_bar) {
    const int mx = height > width ? height : width, mn = height > width ? width : height;
    if ((double)mx / mn > 200.0) return false;
    h_bar = (int)std::nearbyint((double)height / factor) * factor;
    w_bar = (int)std::nearbyint((double)width / factor) * factor;
    if ((long long)h_bar * w_bar > max_pixels) {
        const double beta = std::sqrt(((double)height * width) / (double)max_pixels);
        h_bar = (int)std::floor((double)height / beta / factor) * factor;
        w_bar = (int)std::floor((double)width / beta / factor) * factor;
        if (h_bar < factor) h_bar = factor;
        if (w_bar < factor) w_bar = factor;
    } else if ((long long)h_bar * w_bar < min_pixels) {
        const double beta = std::sqrt((double)min_pixels / ((double)height * width));
        h_bar = (int)std::ceil((double)height * beta / factor) * factor;
        w_bar = (int)std::ceil((double)width * beta / factor) * factor;
    }
    return true;
}

// Qwen still images: smart_resize, plain bicubic resize of the whole image (no canvas, no crop), patch-row output; image i's rows
// start at sum_{j<i} gh_j * gw_j. grid_hw[i] = (gh, gw) in patches.
inline bool pre_plan_qwen(const int32_t* heights, const int32_t* widths, const int64_t* offsets, int n, int patch, int merge, int temporal,
                          long long min_pixels, long long max_pixels, PrePlan& plan, int32_t* grid_hw, long long* total_rows) {
    std::vector<PreTableKey> cache;
    cache.reserve(n);
    plan.images.resize(n);
    long long rows = 0;
    const long long row_elems = 3LL * temporal * patch * patch;
    for (int i = 0; i < n; ++i) {
        const int h = heights[i], w = widths[i];
        if (h <= 0 || w <= 0) { plan.error = "image with non-positive size"; return false; }
        int rh, rw;
        if (!pre_smart_resize(h, w, patch * merge, min_pixels, max_pixels, rh, rw)) {
            plan.error = "absolute aspect ratio must be smaller than 200";
            return false;
        }
        PreImage& im = plan.images[i];
        im.src_off = offsets[i]; im.h = h; im.w = w;
        im.canvas_h = h; im.canvas_w = w; im.paste_x = im.paste_y = 0;
        const PreTableKey* k = pre_tables(plan, cache, h, w, rh, rw, 0, 0, rh, rw);
        if (!k) return false;
        pre_finish_image(plan, im, *k);
        im.out_off = rows * row_elems;
        if (grid_hw) { grid_hw[2 * i] = rh / patch; grid_hw[2 * i + 1] = rw / patch; }
        rows += (long long)(rh / patch) * (rw / patch);
    }
    if (total_rows) *total_rows = rows;
    return true;
}

}  // namespace vqa
