// Connected-component labelling of binarised maps on the device (first piece).
//
//   replaces: label_layer (interpreter/interpreter.py:16-22): `ndimage.label(layer > np.mean(layer))`, the step every
//   crop stage starts with (CropAndRotateParagraphs :437-447, CropRotateAndZoomLines :421-470, LabelChar) -- the
//   reference runs it on the host with SciPy after moving each predicted map off the device.  Input here is the uint8
//   mask the device already holds (uocr_threshold_mask); output is the int32 label map with SciPy's numbering and the
//   component count per image, so only the label map has to cross the host link.
//
// Semantics (bit-exact with scipy.ndimage.label, default structure, on a (1, H, W, 1) array): foreground = pixels
// STRICTLY above the image's mean (for a 0/1 mask: the ones, unless the whole image is ones -- then nothing, as in the
// reference); 4-neighbourhood; labels 1..count in raster order of each component's first pixel.
//
// Algorithm: union-find over linear pixel indices whose root is always a component's SMALLEST index, so that ranking
// the roots by index gives SciPy's numbering.
//   1. mean      per-image sum of the mask (integer atomics: exact)
//   2. runs      one CTA per image row: parent[p] = first pixel of p's horizontal run (block-wide prefix max)
//   3. merge     a run is joined to the run above wherever the pixel above is foreground (once per maximal contact):
//                lock-free union by atomicMin on the larger root
//   4. flatten   parent[p] = root(p); roots are flagged; per-chunk root counts
//   5. rank      exclusive scan of the chunk counts per image (one CTA per image), then labels[root] = rank + 1
//   6. paint     labels[p] = labels[root(p)]     (in place: only root entries are read)
#include "common.cuh"

namespace uocr {
namespace {

constexpr int LB_CHUNK = 1024;           // pixels per ranking chunk

__global__ void __launch_bounds__(256) label_sum_kernel(const uint8_t* __restrict__ mask, int64_t px_per_image,
                                                        unsigned long long* __restrict__ sums) {
    const int64_t n = blockIdx.y;
    const uint8_t* m = mask + n * px_per_image;
    unsigned int acc = 0;
    for (int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x; i < px_per_image; i += (int64_t)gridDim.x * 256) acc += m[i];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    if ((threadIdx.x & 31) == 0 && acc) atomicAdd(&sums[n], (unsigned long long)acc);
}

// fg(p) = mask[p] * P > sum  (exact form of mask[p] > sum / P)
__device__ __forceinline__ bool lb_fg(uint8_t v, unsigned long long sum, unsigned long long px) {
    return (unsigned long long)v * px > sum;
}

// one CTA per (image, row): parent[p] = index (inside the image) of the first pixel of p's run, -1 for background
__global__ void __launch_bounds__(256) label_runs_kernel(const uint8_t* __restrict__ mask, int32_t* __restrict__ parent,
                                                         int h, int w, const unsigned long long* __restrict__ sums) {
    __shared__ int warp_max[8];
    __shared__ int carry;
    const int row = blockIdx.x, n = blockIdx.y;
    const int64_t base = ((int64_t)n * h + row) * w;
    const unsigned long long sum = sums[n], px = (unsigned long long)h * w;
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    if (threadIdx.x == 0) carry = -1;
    __syncthreads();
    for (int c0 = 0; c0 < w; c0 += 256) {
        const int c = c0 + threadIdx.x;
        const bool fg = c < w && lb_fg(mask[base + c], sum, px);
        const bool left = c > 0 && c < w && lb_fg(mask[base + c - 1], sum, px);
        // start of a run at c -> candidate c, else -1; the run start of c is the running maximum of the candidates
        int v = (fg && !left) ? c : -1;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int t = __shfl_up_sync(0xffffffffu, v, o);
            if (lane >= o) v = max(v, t);
        }
        if (lane == 31) warp_max[wid] = v;
        __syncthreads();
        int pre = carry;
        for (int k = 0; k < wid; ++k) pre = max(pre, warp_max[k]);
        v = max(v, pre);
        if (c < w) parent[base + c] = fg ? row * w + v : -1;
        __syncthreads();
        if (threadIdx.x == 255) carry = v;
        __syncthreads();
    }
}

__device__ __forceinline__ int32_t lb_find(int32_t* __restrict__ parent, int32_t p) {
    int32_t r = parent[p];
    while (r != p) {
        const int32_t g = parent[r];
        if (g != r) parent[p] = g;           // path halving (benign race: parents only ever decrease towards the root)
        p = r;
        r = g;
    }
    return p;
}

// read-only walk to the root: used where other threads store final roots concurrently (a path-halving store of a
// stale ancestor could overwrite one)
__device__ __forceinline__ int32_t lb_root(const int32_t* __restrict__ parent, int32_t p) {
    int32_t r = parent[p];
    while (r != p) {
        p = r;
        r = parent[r];
    }
    return p;
}

__device__ __forceinline__ void lb_union(int32_t* __restrict__ parent, int32_t a, int32_t b) {
    while (true) {
        a = lb_find(parent, a);
        b = lb_find(parent, b);
        if (a == b) return;
        if (a < b) { const int32_t t = a; a = b; b = t; }       // a > b: hang the larger root under the smaller
        const int32_t old = atomicMin(&parent[a], b);
        if (old == a) return;
        a = old;                                                 // somebody re-parented a meanwhile: retry from there
    }
}

// join vertically adjacent runs: at every pixel whose upper neighbour is foreground and that is the first such pixel
// of the contact (its left neighbour is not part of the same contact)
__global__ void __launch_bounds__(256) label_merge_kernel(int32_t* __restrict__ parent, int h, int w, int64_t total) {
    const int64_t px = (int64_t)h * w;
    for (int64_t g = (int64_t)blockIdx.x * 256 + threadIdx.x; g < total; g += (int64_t)gridDim.x * 256) {
        const int64_t n = g / px;
        const int32_t p = (int32_t)(g - n * px);
        int32_t* par = parent + n * px;
        if (p < w || par[p] < 0 || par[p - w] < 0) continue;
        const int c = p % w;
        if (c > 0 && par[p - 1] >= 0 && par[p - 1 - w] >= 0) continue;      // same contact as the pixel to the left
        lb_union(par, p, p - w);
    }
}

// parent[p] = root(p); chunk_roots[chunk] = number of roots inside the chunk
__global__ void __launch_bounds__(256) label_flatten_kernel(int32_t* __restrict__ parent, int64_t px, int chunks,
                                                            int32_t* __restrict__ chunk_roots) {
    __shared__ int cnt;
    const int chunk = blockIdx.x, n = blockIdx.y;
    int32_t* par = parent + (int64_t)n * px;
    if (threadIdx.x == 0) cnt = 0;
    __syncthreads();
    int mine = 0;
    for (int k = threadIdx.x; k < LB_CHUNK; k += 256) {
        const int64_t p = (int64_t)chunk * LB_CHUNK + k;
        if (p >= px || par[p] < 0) continue;
        const int32_t r = lb_root(par, (int32_t)p);
        par[p] = r;                          // fully flattened: the paint pass reads parent[p] as THE root
        mine += r == (int32_t)p;
    }
    if (mine) atomicAdd(&cnt, mine);
    __syncthreads();
    if (threadIdx.x == 0) chunk_roots[(int64_t)n * chunks + chunk] = cnt;
}

// one CTA per image: exclusive scan of its chunk counts (in place) and the image's component count
__global__ void __launch_bounds__(1024) label_scan_kernel(int32_t* __restrict__ chunk_roots, int chunks,
                                                          int32_t* __restrict__ counts) {
    __shared__ int warp_sum[32];
    __shared__ int carry;
    int32_t* cr = chunk_roots + (int64_t)blockIdx.x * chunks;
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    if (threadIdx.x == 0) carry = 0;
    __syncthreads();
    for (int c0 = 0; c0 < chunks; c0 += 1024) {
        const int c = c0 + threadIdx.x;
        const int own = c < chunks ? cr[c] : 0;
        int v = own;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int t = __shfl_up_sync(0xffffffffu, v, o);
            if (lane >= o) v += t;
        }
        if (lane == 31) warp_sum[wid] = v;
        __syncthreads();
        int pre = carry;
        for (int k = 0; k < wid; ++k) pre += warp_sum[k];
        if (c < chunks) cr[c] = pre + v - own;
        __syncthreads();
        if (threadIdx.x == 1023) carry = pre + v;
        __syncthreads();
    }
    if (threadIdx.x == 0) counts[blockIdx.x] = carry;
}

// labels[root] = 1 + (roots before it in raster order): chunk offset + rank inside the chunk (one warp-ballot pass)
__global__ void __launch_bounds__(256) label_rank_kernel(const int32_t* __restrict__ parent, int64_t px, int chunks,
                                                         const int32_t* __restrict__ chunk_roots, int32_t* __restrict__ labels) {
    __shared__ int warp_cnt[8];
    __shared__ int carry;
    const int chunk = blockIdx.x, n = blockIdx.y;
    const int32_t* par = parent + (int64_t)n * px;
    int32_t* lab = labels + (int64_t)n * px;
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    if (threadIdx.x == 0) carry = chunk_roots[(int64_t)n * chunks + chunk];
    __syncthreads();
    for (int k0 = 0; k0 < LB_CHUNK; k0 += 256) {
        const int64_t p = (int64_t)chunk * LB_CHUNK + k0 + threadIdx.x;
        const bool root = p < px && par[p] == (int32_t)p;
        const unsigned ballot = __ballot_sync(0xffffffffu, root);
        if (lane == 0) warp_cnt[wid] = __popc(ballot);
        __syncthreads();
        int pre = carry;
        for (int k = 0; k < wid; ++k) pre += warp_cnt[k];
        if (root) lab[p] = pre + __popc(ballot & ((1u << lane) - 1u)) + 1;
        __syncthreads();
        if (threadIdx.x == 0) {
            int tot = 0;
            for (int k = 0; k < 8; ++k) tot += warp_cnt[k];
            carry += tot;
        }
        __syncthreads();
    }
}

__global__ void __launch_bounds__(256) label_paint_kernel(const int32_t* __restrict__ parent, int32_t* __restrict__ labels,
                                                          int64_t px, int64_t total) {
    for (int64_t g = (int64_t)blockIdx.x * 256 + threadIdx.x; g < total; g += (int64_t)gridDim.x * 256) {
        const int64_t n = g / px;
        const int32_t r = parent[g];
        if (r < 0) labels[g] = 0;
        else if (r != (int32_t)(g - n * px)) labels[g] = labels[n * px + r];     // root entries are final already
    }
}

// ---- per-component statistics: what the crop stages take from each object mask (`labels == l`): its bounding box
// (ndimage.find_objects, interpreter.py:125-148, 230, 303, 341, 496-497) and its centre of mass
// (ndimage.center_of_mass, :36-38, 142-144).  stats[n][l - 1][0..6] = count, sum_y, sum_x, y_min, y_max, x_min, x_max.
// Lanes of a warp that hold the same label are combined first (__match_any_sync): one set of atomics per label and warp.
constexpr int LB_STATS = 7;

__global__ void __launch_bounds__(256) label_stats_init_kernel(long long* __restrict__ stats, int64_t entries) {
    for (int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x; i < entries; i += (int64_t)gridDim.x * 256) {
        long long* s = stats + i * LB_STATS;
        s[0] = 0; s[1] = 0; s[2] = 0;
        s[3] = 0x7fffffffffffffffll; s[4] = -1; s[5] = 0x7fffffffffffffffll; s[6] = -1;
    }
}

__global__ void __launch_bounds__(256) label_stats_kernel(const int32_t* __restrict__ labels, long long* __restrict__ stats,
                                                          int h, int w, int64_t max_labels, int64_t total) {
    const int64_t px = (int64_t)h * w;
    const int lane = threadIdx.x & 31;
    // whole warps iterate together (the loop bound is rounded up per warp), idle lanes carry label 0
    for (int64_t g0 = ((int64_t)blockIdx.x * 256 + (threadIdx.x & ~31)); g0 < total; g0 += (int64_t)gridDim.x * 256) {
        const int64_t g = g0 + lane;
        int32_t l = 0;
        int y = 0, x = 0;
        int64_t n = 0;
        if (g < total) {
            l = labels[g];
            n = g / px;
            const int64_t p = g - n * px;
            y = (int)(p / w);
            x = (int)(p - (int64_t)y * w);
        }
        if (l > max_labels) l = 0;
        // key = (image, label): lanes of one warp can straddle two images
        const long long key = l > 0 ? (n << 32) | (long long)l : -1 - lane;
        const unsigned peers = __match_any_sync(0xffffffffu, key);
        if (l <= 0) continue;
        const int leader = __ffs(peers) - 1;
        long long cnt = 0, sy = 0, sx = 0;
        int y0 = y, y1 = y, x0 = x, x1 = x;
        for (unsigned m = peers; m; m &= m - 1) {
            const int src = __ffs(m) - 1;
            const int oy = __shfl_sync(peers, y, src), ox = __shfl_sync(peers, x, src);
            ++cnt; sy += oy; sx += ox;
            y0 = min(y0, oy); y1 = max(y1, oy); x0 = min(x0, ox); x1 = max(x1, ox);
        }
        if (lane == leader) {
            long long* s = stats + (n * max_labels + (l - 1)) * LB_STATS;
            atomicAdd(reinterpret_cast<unsigned long long*>(s + 0), (unsigned long long)cnt);
            atomicAdd(reinterpret_cast<unsigned long long*>(s + 1), (unsigned long long)sy);
            atomicAdd(reinterpret_cast<unsigned long long*>(s + 2), (unsigned long long)sx);
            atomicMin(s + 3, (long long)y0);
            atomicMax(s + 4, (long long)y1);
            atomicMin(s + 5, (long long)x0);
            atomicMax(s + 6, (long long)x1);
        }
    }
}

}  // namespace
}  // namespace uocr

using namespace uocr;

extern "C" {

int uocr_label_components_workspace(int64_t n, int64_t h, int64_t w, size_t* bytes) {
    UOCR_REQUIRE(bytes && n > 0 && h > 0 && w > 0, "bad argument");
    UOCR_REQUIRE(h * w < (1ll << 31), "image too large for 32-bit pixel indices");
    const int64_t px = h * w, chunks = ceil_div(px, LB_CHUNK);
    // parent (int32 per pixel) | chunk root counts (int32 per chunk) | per-image sums (uint64), 16-byte aligned parts
    *bytes = (size_t)(((n * px * 4 + 15) & ~15ll) + ((n * chunks * 4 + 15) & ~15ll) + n * 8);
    return UOCR_OK;
}

int uocr_label_components(const uint8_t* mask, int32_t* labels, int32_t* counts, int64_t n, int64_t h, int64_t w,
                          void* workspace, void* stream) {
    UOCR_REQUIRE(mask && labels && counts && workspace, "NULL pointer");
    UOCR_REQUIRE(n > 0 && h > 0 && w > 0 && n <= 65535 && h <= 0x7fffffff, "bad dimension");
    UOCR_REQUIRE(h * w < (1ll << 31), "image too large for 32-bit pixel indices");
    cudaStream_t st = as_stream(stream);
    const int64_t px = h * w, chunks = ceil_div(px, LB_CHUNK), total = n * px;
    UOCR_REQUIRE(chunks <= 0x7fffffff && h <= 0x7fffffff, "image too large");
    char* ws = static_cast<char*>(workspace);
    int32_t* parent = reinterpret_cast<int32_t*>(ws);
    int32_t* chunk_roots = reinterpret_cast<int32_t*>(ws + ((n * px * 4 + 15) & ~15ll));
    unsigned long long* sums = reinterpret_cast<unsigned long long*>(ws + ((n * px * 4 + 15) & ~15ll) + ((n * chunks * 4 + 15) & ~15ll));
    UOCR_CUDA(cudaMemsetAsync(sums, 0, (size_t)n * 8, st));
    const int sum_blocks = (int)(ceil_div(px, 256 * 16) < 1024 ? ceil_div(px, 256 * 16) : 1024);
    label_sum_kernel<<<dim3((unsigned)sum_blocks, (unsigned)n), 256, 0, st>>>(mask, px, sums);
    UOCR_LAUNCHED("label_sum");
    label_runs_kernel<<<dim3((unsigned)h, (unsigned)n), 256, 0, st>>>(mask, parent, (int)h, (int)w, sums);
    UOCR_LAUNCHED("label_runs");
    const int flat_blocks = (int)(ceil_div(total, 256) < 148 * 32 ? ceil_div(total, 256) : 148 * 32);
    label_merge_kernel<<<flat_blocks, 256, 0, st>>>(parent, (int)h, (int)w, total);
    UOCR_LAUNCHED("label_merge");
    label_flatten_kernel<<<dim3((unsigned)chunks, (unsigned)n), 256, 0, st>>>(parent, px, (int)chunks, chunk_roots);
    UOCR_LAUNCHED("label_flatten");
    label_scan_kernel<<<(unsigned)n, 1024, 0, st>>>(chunk_roots, (int)chunks, counts);
    UOCR_LAUNCHED("label_scan");
    label_rank_kernel<<<dim3((unsigned)chunks, (unsigned)n), 256, 0, st>>>(parent, px, (int)chunks, chunk_roots, labels);
    UOCR_LAUNCHED("label_rank");
    label_paint_kernel<<<flat_blocks, 256, 0, st>>>(parent, labels, px, total);
    UOCR_LAUNCHED("label_paint");
    return UOCR_OK;
}

int uocr_label_stats(const int32_t* labels, int64_t* stats, int64_t n, int64_t h, int64_t w, int64_t max_labels,
                     void* stream) {
    UOCR_REQUIRE(labels && stats, "NULL pointer");
    UOCR_REQUIRE(n > 0 && h > 0 && w > 0 && max_labels > 0 && h <= 0x7fffffff && w <= 0x7fffffff, "bad dimension");
    UOCR_REQUIRE(h * w < (1ll << 31) && max_labels < (1ll << 31), "image too large for 32-bit pixel indices");
    cudaStream_t st = as_stream(stream);
    const int64_t entries = n * max_labels, total = n * h * w;
    const int init_blocks = (int)(ceil_div(entries, 256) < 148 * 8 ? ceil_div(entries, 256) : 148 * 8);
    label_stats_init_kernel<<<init_blocks, 256, 0, st>>>(reinterpret_cast<long long*>(stats), entries);
    UOCR_LAUNCHED("label_stats_init");
    const int blocks = (int)(ceil_div(total, 256) < 148 * 32 ? ceil_div(total, 256) : 148 * 32);
    label_stats_kernel<<<blocks, 256, 0, st>>>(labels, reinterpret_cast<long long*>(stats), (int)h, (int)w, max_labels, total);
    UOCR_LAUNCHED("label_stats");
    return UOCR_OK;
}

}  // extern "C"
