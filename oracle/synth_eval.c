/* synth_eval.c — TEST INFRASTRUCTURE ONLY (oracle).  Not part of the product path.
 *
 * A deterministic stand-in for the network with the leaf contract of the substitute Evaluator
 * (maru_b200/dropin/states/Evaluator.cpp: 448-byte maru_state in, maru_leaf out), so that the
 * UNMODIFIED reference search (deepgo::Player / Node, compiled into
 * oracle/_ref/libmaru_ref_search_synth.so) and the native search (maru_b200/csrc/search.h,
 * maru_selfplay_create_fn) can be driven by the SAME evaluator on the CPU and compared move by move
 * (tests/test_search.py).  Priors and value are a hash of the whole record: every position gets its
 * own peaked policy and its own value, nothing about Go is modelled.
 */
#include <math.h>
#include <stdint.h>
#include <string.h>

#include "maru_b200.h"

static uint64_t mix(uint64_t x) {
  x ^= x >> 33;
  x *= 0xff51afd7ed558ccdull;
  x ^= x >> 33;
  x *= 0xc4ceb9fe1a85ec53ull;
  x ^= x >> 33;
  return x;
}

static void synth_one(const maru_state* s, maru_leaf* out) {
  memset(out, 0, sizeof(*out));
  uint64_t h = 0x9e3779b97f4a7c15ull;
  const int cells = s->width * s->height;
  for (int i = 0; i < cells; i++) h = mix(h ^ ((uint64_t)s->cells[i] + 131u * (uint64_t)i));   /* colour, liberties AND ladder flag */
  h = mix(h ^ (uint64_t)(uint8_t)s->color);
  h = mix(h ^ ((uint64_t)s->ko_x << 8 | s->ko_y));
  for (int i = 0; i < 12; i++) h = mix(h ^ ((uint64_t)((const uint8_t*)s->hist)[i] << (i & 7)));   /* move history */
  double w[361], sum = 0.0;
  for (int i = 0; i < cells; i++) {
    const double u = (double)(mix(h ^ (0x1234567ull * (uint64_t)(i + 1))) >> 11) / 9007199254740992.0;
    w[i] = exp(4.0 * u);                       /* peaked: a few moves carry most of the mass */
    sum += w[i];
  }
  const int masked = (s->flags & MARU_STATE_HAS_MASK) != 0;
  for (int i = 0; i < cells; i++) {
    const int keep = !masked || ((s->move_mask[i >> 3] >> (i & 7)) & 1u);
    out->prior[i] = keep ? (float)(w[i] / sum) : 0.0f;
  }
  const double v = (double)(mix(h ^ 0xabcdefull) >> 11) / 9007199254740992.0;
  out->value = (float)(0.1 + 0.8 * v);         /* win probability of the side to move */
  out->score[0] = 0.0f;
  out->score[1] = 0.0f;
}

/* maru_leaf_fn (include/maru_b200.h) */
int oracle_synth_leaves(const maru_state* s, maru_leaf* out, int n, void* user) {
  (void)user;
  for (int i = 0; i < n; i++) synth_one(&s[i], &out[i]);
  return 0;
}
