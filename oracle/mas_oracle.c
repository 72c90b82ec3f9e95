/*
 * oracle/mas_oracle.c -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * Plain-C restatement of the reference's Monotonic Alignment Search
 * (reference: monotonic_align/core.pyx:7-42).  Only tests/, __graft_entry__.smoke()
 * and bench.py's cpu_baseline / --impl reference legs may load this file's .so;
 * the product path (vits_b200/) never does.
 *
 * Parity pin: this restatement is checked bit-for-bit against the reference's own
 * compiled Cython (oracle/_ref, built by oracle/Makefile from the sources where they
 * lie under /root/reference) and against tests/golden/ fixtures produced by that
 * compiled reference (tests/golden/make_golden.py).
 *
 * Build:  gcc -O2 -fPIC -shared -ffp-contract=off -fno-fast-math (see oracle/Makefile)
 *
 * Semantics restated (all from core.pyx):
 *   - forward sweep over rows y, band  max(0, t_x+y-t_y) <= x < min(t_x, y+1)   (:15-16)
 *   - "stay" candidate is value[y-1][x], replaced by -1e9 on the diagonal x==y    (:17-20)
 *   - "step" candidate is value[y-1][x-1]; at x==0 it is 0 for y==0 else -1e9     (:21-27)
 *   - value[y][x] += (stay > step) ? stay : step     (Cython's max(a,b) on floats) (:28)
 *   - backtrack from (t_y-1, t_x-1); step left when index==y or
 *     value[y-1][index] < value[y-1][index-1]  (strict: ties stay)               (:30-33)
 */
#include <stdint.h>
#include <stddef.h>

#define MAS_NEG_SENTINEL (-1e9f)

/* One utterance.  `value` is clobbered (the reference accumulates in place);
 * `path` must be pre-zeroed by the caller (reference: __init__.py:15).
 * row_stride = padded T_x of the slab both arrays live in.                    */
static void mas_oracle_one(int32_t *path, float *value, int t_y, int t_x, ptrdiff_t row_stride)
{
  for (int y = 0; y < t_y; ++y) {
    int lo = t_x + y - t_y;
    if (lo < 0) lo = 0;
    int hi = (y + 1 < t_x) ? y + 1 : t_x;              /* exclusive */
    float *cur = value + (ptrdiff_t)y * row_stride;
    const float *up = cur - row_stride;                /* only dereferenced when y > 0 */
    for (int x = lo; x < hi; ++x) {
      float stay = (x == y) ? MAS_NEG_SENTINEL : up[x];
      float step;
      if (x == 0)
        step = (y == 0) ? 0.0f : MAS_NEG_SENTINEL;
      else
        step = up[x - 1];
      float best = (stay > step) ? stay : step;
      cur[x] = cur[x] + best;
    }
  }

  int index = t_x - 1;
  for (int y = t_y - 1; y >= 0; --y) {
    path[(ptrdiff_t)y * row_stride + index] = 1;
    if (index != 0) {
      const float *up = value + (ptrdiff_t)(y - 1) * row_stride;
      if (index == y || up[index] < up[index - 1])
        --index;
    }
  }
}

/* Batch entry, same argument meaning as the reference's maximum_path_c
 * (core.pyx:38): paths int32 [B,T_y,T_x] pre-zeroed, values float32 [B,T_y,T_x]
 * (clobbered), t_ys/t_xs int32 [B].  Serial over the batch, like the shipped build. */
void mas_oracle_maximum_path(int32_t *paths, float *values, const int32_t *t_ys,
                             const int32_t *t_xs, int B, int T_y, int T_x)
{
  const ptrdiff_t plane = (ptrdiff_t)T_y * T_x;
  for (int b = 0; b < B; ++b)
    mas_oracle_one(paths + b * plane, values + b * plane, t_ys[b], t_xs[b], T_x);
}

/* Convenience for tests: per-frame text index (the column holding the 1), -1 on
 * padded frames.  Derived from a finished `paths` tensor.                        */
void mas_oracle_path_to_index(const int32_t *paths, int32_t *index, int B, int T_y, int T_x)
{
  for (int b = 0; b < B; ++b)
    for (int y = 0; y < T_y; ++y) {
      const int32_t *row = paths + ((ptrdiff_t)b * T_y + y) * T_x;
      int32_t found = -1;
      for (int x = 0; x < T_x; ++x)
        if (row[x]) { found = x; break; }
      index[(ptrdiff_t)b * T_y + y] = found;
    }
}
