/* TEST INFRASTRUCTURE ONLY — plain-C restatement of farthest point sampling as RaLD uses it.
 *
 * The reference calls torch_cluster.fps (third party, torch_cluster==1.6.3+pt25cu124 per requirements.txt:156; its
 * source is not in the reference tree) at model/models_ae.py:243 and :368 as fps(pos[B*N,3], batch, ratio=M/N).
 * Published algorithm of torch_cluster 1.6.x: per cloud, m = ceil(ratio*n) picks; first pick = start index (random in
 * upstream's default random_start=True; FIXED TO 0 here = random_start=False); dist[i] = |p_i - p_start|^2; then
 * repeat: next = argmax(dist); dist[i] = min(dist[i], |p_i - p_next|^2). Output = picked indices in pick order.
 * No reference test or golden vector pins the result: PARITY UNPINNED for FPS. Conventions made
 * explicit here and shared with the CUDA kernel: squared distance = (dx*dx + dy*dy) + dz*dz with every fp32
 * operation rounded separately (build with -ffp-contract=off), argmax ties -> lowest index.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>

int fps_oracle(const float* pts, int B, int N, int M, int64_t* out) {
  float* dist = (float*)malloc(sizeof(float) * (size_t)N);
  if (!dist) return -1;
  for (int b = 0; b < B; ++b) {
    const float* p = pts + (size_t)b * N * 3;
    for (int i = 0; i < N; ++i) dist[i] = INFINITY;
    int cur = 0;
    for (int m = 0; m < M; ++m) {
      out[(size_t)b * M + m] = cur;
      const float cx = p[3 * cur], cy = p[3 * cur + 1], cz = p[3 * cur + 2];
      float best = -1.0f;
      int besti = 0;
      for (int i = 0; i < N; ++i) {
        const float dx = p[3 * i] - cx, dy = p[3 * i + 1] - cy, dz = p[3 * i + 2] - cz;
        const float xx = dx * dx, yy = dy * dy, zz = dz * dz;
        const float s = xx + yy;
        const float d2 = s + zz;
        const float d = dist[i] < d2 ? dist[i] : d2;
        dist[i] = d;
        if (d > best) { best = d; besti = i; }
      }
      cur = besti;
    }
  }
  free(dist);
  return 0;
}
