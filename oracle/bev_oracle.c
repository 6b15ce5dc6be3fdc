/*
 * Plain-C oracle for the LiDAR -> BEV projection.  TEST INFRASTRUCTURE ONLY:
 * loaded by tests/ and by bench.py's cpu_baseline leg, never by the product.
 *
 * Restates, one IEEE fp32 rounding per reference op:
 *   reference src/models/lidar_encoder.py:42-55  (normalise + closed-range mask)
 *   reference src/models/lidar_encoder.py:69-71  (scale by W-1 / H-1, truncate, clamp)
 *   reference src/models/lidar_encoder.py:77-96  (flat index, amax scatter into zeros)
 *
 * Build: see oracle/Makefile (gcc -O2 -ffp-contract=off, no fast-math).
 * Pinned against the reference through tests/test_oracle_vs_reference.py and
 * the SURVEY.md section-4 known-answer vector.
 */
#include <math.h>
#include <stdint.h>
#include <string.h>

/* cell[i] = row*W + col, or -1 when the point is outside the closed range. */
void bevo_cells(const float *points, int64_t n_points, int point_stride,
                float x0, float xspan, float y0, float yspan, int H, int W,
                int32_t *cell)
{
    const float sx = (float)(W - 1), sy = (float)(H - 1);
    for (int64_t i = 0; i < n_points; ++i) {
        const float x = points[i * point_stride + 0];
        const float y = points[i * point_stride + 1];
        volatile float dx = x - x0;          /* :47 subtraction, rounded to fp32 */
        volatile float dy = y - y0;
        volatile float xn = dx / xspan;      /* :47 true division                 */
        volatile float yn = dy / yspan;
        const int ok = (xn >= 0.0f) && (xn <= 1.0f) && (yn >= 0.0f) && (yn <= 1.0f);
        if (!ok) { cell[i] = -1; continue; }
        volatile float gx = xn * sx;         /* :69 multiply, then .long() truncation */
        volatile float gy = yn * sy;
        int col = (int)gx, row = (int)gy;
        if (col < 0) col = 0; if (col > W - 1) col = W - 1;   /* :70-71 */
        if (row < 0) row = 0; if (row > H - 1) row = H - 1;
        cell[i] = row * W + col;
    }
}

/* occupancy[b*HW + c] = number of points of frame b in cell c. */
void bevo_occupancy(const int32_t *cell, int B, int64_t N, int HW, int32_t *occ)
{
    memset(occ, 0, sizeof(int32_t) * (size_t)B * HW);
    for (int b = 0; b < B; ++b)
        for (int64_t i = 0; i < N; ++i) {
            int32_t c = cell[(int64_t)b * N + i];
            if (c >= 0) occ[(int64_t)b * HW + c]++;
        }
}

/* grid[b,c,:] = max over the points of cell c of feats[b,i,:]; empty cells 0.
 * (amax with include_self=False into a zero tensor, :85-96).  feats is
 * point-major [B,N,C]; grid is cell-major [B,HW,C] which is the memory order of
 * the tensor the reference returns (:99 permuted view). */
void bevo_scatter_max(const float *feats, const int32_t *cell, int B, int64_t N, int C,
                      int HW, float *grid, uint8_t *seen /* [B*HW] scratch */)
{
    memset(grid, 0, sizeof(float) * (size_t)B * HW * C);
    memset(seen, 0, (size_t)B * HW);
    for (int b = 0; b < B; ++b)
        for (int64_t i = 0; i < N; ++i) {
            int32_t c = cell[(int64_t)b * N + i];
            if (c < 0) continue;
            float *g = grid + ((int64_t)b * HW + c) * C;
            const float *f = feats + ((int64_t)b * N + i) * C;
            if (!seen[(int64_t)b * HW + c]) {
                seen[(int64_t)b * HW + c] = 1;
                memcpy(g, f, sizeof(float) * C);
            } else {
                for (int k = 0; k < C; ++k) g[k] = f[k] > g[k] ? f[k] : g[k];
            }
        }
}
