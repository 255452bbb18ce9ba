// Host-side (native C++) glue between the splat and the autoregressive sampler: replaces the Python / OpenCV /
// Cython work of ZbufferModelPts.get_masks_for_batch (reference models/z_buffermodel.py:641-701):
//   AvgPool2d(8) of the background / foreground masks truncated to uint8 (:646-647,668-669),
//   cv2.distanceTransform(DIST_L2, 5) of both (:673-674) -- OpenCV's two-pass 5x5 chamfer transform,
//   distances = int(fd - bd) (:675), the frontier-heap generation order of get_custom_order.pyx:4-124 and the three
//   locally-masked-convolution masks of models/lmconv/masking.py:287-370, emitted as nine-bit words per cell instead
//   of the reference's (B*513|160|80, 9, 1024) float tensors (27.8 MB/image -> 6 KB/image).
// All pointers of this file's entry point are HOST pointers: the input is the 64 KB/image background mask the
// splat produced, the outputs are a few KB per image for the sampler kernel.
#include <math.h>

#include <algorithm>
#include <queue>
#include <tuple>
#include <stdlib.h>

#include <thread>
#include <vector>

#include "common.cuh"
#include "hostpool.cuh"

namespace ps {

static const int G = 32;  // latent grid side (obs = [3, 32, 32], z_buffermodel.py:79)

// cv2.distanceTransform(src, DIST_L2, 5): two-pass 5x5 chamfer transform, weights (1, 1.4, 2.1969), distance to
// the nearest zero cell.  Arithmetic follows the OpenCV build in this image (4.13: float accumulation, FLT_MAX where
// the image has no zero cell); each candidate is one float add, so the result is order independent and bit-exact.
// OpenCV 4.2 (the reference's pin, docs/INSTALL.md:60) accumulated in 16.16 fixed point; the two differ by < 1e-4,
// which matters only through the reference's astype(int) truncation at exact-integer distances.
static void chamfer5x5(const uint8_t* src, int rows, int cols, float* dist) {
  const int B = 2;
  const float HV = 1.0f, DIAG = 1.4f, LONG = 2.1969f;
  const float DMAX = 3.402823466e+38f;
  const int step = cols + 2 * B;
  std::vector<float> temp((size_t)(rows + 2 * B) * step, DMAX);
  for (int i = 0; i < rows; ++i) {
    float* t = temp.data() + (size_t)(i + B) * step + B;
    for (int j = 0; j < cols; ++j) {
      if (!src[i * cols + j]) {
        t[j] = 0.0f;
      } else {
        float t0 = t[j - step * 2 - 1] + LONG;
        t0 = std::min(t0, t[j - step * 2 + 1] + LONG);
        t0 = std::min(t0, t[j - step - 2] + LONG);
        t0 = std::min(t0, t[j - step - 1] + DIAG);
        t0 = std::min(t0, t[j - step] + HV);
        t0 = std::min(t0, t[j - step + 1] + DIAG);
        t0 = std::min(t0, t[j - step + 2] + LONG);
        t0 = std::min(t0, t[j - 1] + HV);
        t[j] = t0;
      }
    }
  }
  for (int i = rows - 1; i >= 0; --i) {
    float* t = temp.data() + (size_t)(i + B) * step + B;
    for (int j = cols - 1; j >= 0; --j) {
      float t0 = t[j];
      if (t0 > HV) {
        t0 = std::min(t0, t[j + step * 2 + 1] + LONG);
        t0 = std::min(t0, t[j + step * 2 - 1] + LONG);
        t0 = std::min(t0, t[j + step + 2] + LONG);
        t0 = std::min(t0, t[j + step + 1] + DIAG);
        t0 = std::min(t0, t[j + step] + HV);
        t0 = std::min(t0, t[j + step - 1] + DIAG);
        t0 = std::min(t0, t[j + step - 2] + LONG);
        t0 = std::min(t0, t[j + 1] + HV);
        t[j] = t0;
      }
      dist[i * cols + j] = t0;
    }
  }
}

// get_custom_order.pyx:55-82 with a real priority queue: key (-distance, r, c) is a total order, so any heap pops
// the same sequence as Python's heapq over (-distances[r,c], [r,c]).
static void custom_order(const long long* dist, int* order) {
  int am = 0;
  for (int i = 1; i < G * G; ++i)
    if (dist[i] > dist[am]) am = i;  // first maximum, row-major (np.argmax)
  typedef std::tuple<long long, int, int> Key;
  std::priority_queue<Key, std::vector<Key>, std::greater<Key>> heap;
  std::vector<char> seen(G * G, 0);
  int r = am / G, c = am % G, n = 0;
  seen[am] = 1;
  order[n++] = am;
  const int dr[4] = {-1, 1, 0, 0}, dc[4] = {0, 0, -1, 1};  // Up, Down, Left, Right
  while (n < G * G) {
    for (int k = 0; k < 4; ++k) {
      const int rr = r + dr[k], cc = c + dc[k];
      if (rr >= 0 && rr < G && cc >= 0 && cc < G && !seen[rr * G + cc]) {
        seen[rr * G + cc] = 1;
        heap.push(Key(-dist[rr * G + cc], rr, cc));
      }
    }
    const Key k = heap.top();
    heap.pop();
    r = std::get<1>(k);
    c = std::get<2>(k);
    order[n++] = r * G + c;
  }
}

static void mask_words(const int* order, uint16_t* words /* [3][G*G] */) {
  std::vector<int> rank(G * G);
  for (int i = 0; i < G * G; ++i) rank[order[i]] = i;
  const int dil[3] = {1, 1, 2}, centre[3] = {0, 1, 1};
  for (int m = 0; m < 3; ++m)
    for (int r = 0; r < G; ++r)
      for (int c = 0; c < G; ++c) {
        unsigned w = 0;
        for (int t = 0; t < 9; ++t) {
          const int d_r = t / 3 - 1, d_c = t % 3 - 1;
          if (d_r == 0 && d_c == 0) {
            w |= (unsigned)centre[m] << t;
            continue;
          }
          const int rr = r + d_r * dil[m], cc = c + d_c * dil[m];
          if (rr >= 0 && rr < G && cc >= 0 && cc < G && rank[rr * G + cc] < rank[r * G + c]) w |= 1u << t;
        }
        words[m * G * G + r * G + c] = (uint16_t)w;
      }
}

}  // namespace ps

using namespace ps;

extern "C" int ps_lmconv_glue_host(const uint8_t* bg_mask_host, int B, int S, int* dist_host, int* order_host,
                                   uint16_t* words_host, uint8_t* sample_mask_host) {
  PS_CHECK_ARG(bg_mask_host && order_host && words_host && sample_mask_host && B >= 0 && S == 8 * G);
  // images are independent: a few host threads share them (the GPU is waiting for this result)
  auto work = [&](int b0, int b1) {
    std::vector<uint8_t> fg(G * G), bg(G * G);
    std::vector<float> fd(G * G), bd(G * G);
    std::vector<long long> d(G * G);
    for (int b = b0; b < b1; ++b) {
      const uint8_t* m = bg_mask_host + (size_t)b * S * S;
      for (int r = 0; r < G; ++r)
        for (int c = 0; c < G; ++c) {
          // 64 mask bytes (0 / 1) of the cell: eight 8-byte words, byte sums by multiplication
          int cnt = 0;
          for (int y = 0; y < 8; ++y) {
            uint64_t v;
            memcpy(&v, m + (size_t)(r * 8 + y) * S + c * 8, 8);
            v = (v | (v >> 1) | (v >> 2) | (v >> 3) | (v >> 4) | (v >> 5) | (v >> 6) | (v >> 7)) & 0x0101010101010101ull;  // any non-zero byte counts once
            cnt += (int)((v * 0x0101010101010101ull) >> 56);
          }
          bg[r * G + c] = cnt == 64;  // AvgPool2d(8) -> astype(uint8): 1 only when every pixel agrees
          fg[r * G + c] = cnt == 0;
          sample_mask_host[(size_t)b * G * G + r * G + c] = cnt == 64;  // sample.py:29 `== 1`
        }
      chamfer5x5(fg.data(), G, G, fd.data());
      chamfer5x5(bg.data(), G, G, bd.data());
      for (int i = 0; i < G * G; ++i) {
        // a transform with no zero cell saturates as OpenCV 4.2 did ((UINT_MAX - LONG_DIST) / 65536): the FLT_MAX of
        // newer builds does not survive the reference's astype(int)
        const float sat = (float)((4294967295.0 - 143976.0) / 65536.0);
        const double v = (double)std::min(fd[i], sat) - (double)std::min(bd[i], sat);  // float64, z_buffermodel.py:670-675
        d[i] = (long long)v;                             // astype(int) truncates toward zero
        if (dist_host) dist_host[(size_t)b * G * G + i] = (int)d[i];
        d[i] *= 10000;                                   // get_custom_order.pyx:26
      }
      custom_order(d.data(), order_host + (size_t)b * G * G);
      mask_words(order_host + (size_t)b * G * G, words_host + (size_t)b * 3 * G * G);
    }
  };
  unsigned hw = std::thread::hardware_concurrency();
  if (const char* e = getenv("PS_HOST_THREADS")) hw = (unsigned)std::max(1, atoi(e));  // one process per GPU: cores / world
  int nth = (int)std::min<unsigned>(hw ? hw : 1u, 16u);
  nth = std::min(nth, B / 2);  // at least two images per thread
  if (nth <= 1) {
    work(0, B);
  } else {
    HostPool::get().run(nth, [&](int t) { work((int)((long long)B * t / nth), (int)((long long)B * (t + 1) / nth)); });
  }
  return PS_OK;
}
