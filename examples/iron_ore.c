/*
 * quokkasim_examples/src/bin/iron_ore.rs over the C ABI, in plain C99: a user-defined resource (IronOre: five masses,
 * total() = fe + other_elements), IronOreStock -> IronOreProcess -> IronOreStock, quantity Uniform(2, 8), 10 s per
 * batch, step_until(300 s) -- many replications at once.  What a Rust / C / C++ host would emit for that example.
 *
 *   gcc -std=c99 -Iinclude examples/iron_ore.c -o iron_ore -Lquokkasim_b200 -l:libquokka_b200.so \
 *       -Wl,-rpath,$PWD/quokkasim_b200 && ./iron_ore [n_replications]
 */
#include <stdio.h>
#include <stdlib.h>

#include "quokka_b200.h"

#define CHECK(call)                                                      \
  do {                                                                   \
    int rc_ = (call);                                                    \
    if (rc_ != QK_OK) {                                                  \
      fprintf(stderr, "%s -> %d: %s\n", #call, rc_, qk_last_error());    \
      return 1;                                                          \
    }                                                                    \
  } while (0)

enum { FE, OTHER_ELEMENTS, MAGNETITE, HEMATITE, LIMONITE, IRON_ORE_FIELDS };
#define IRON_ORE_TOTAL_MASK ((1u << FE) | (1u << OTHER_ELEMENTS)) /* iron_ore.rs:75-79 */

int main(int argc, char** argv) {
  const unsigned long long n_reps = argc > 1 ? strtoull(argv[1], 0, 10) : 4096ull;
  qk_model* model = 0;
  qk_batch* batch = 0;
  qk_component_desc stock = {QK_VECTOR_STOCK, 0, 0, 0, 0, 0}, process = {QK_VECTOR_PROCESS, 0, 0, 0, 0, 0};
  unsigned s1, p1, s2, d_time, d_qty;
  const double init1[IRON_ORE_FIELDS] = {60.0, 40.0, 10.0, 5.0, 15.0}, init2[IRON_ORE_FIELDS] = {3.0, 2.0, 0.5, 0.25, 0.75};
  qk_dist_desc quantity = {QK_DIST_UNIFORM, 123456789u, {2.0, 8.0, 0.0, 0.0}}; /* df.create(..): seed = stream */
  qk_dist_desc time10 = {QK_DIST_CONSTANT, 0u, {10.0, 0.0, 0.0, 0.0}};
  qk_batch_config cfg;
  double ore[QK_VEC_MAX_DIM];
  double sum[IRON_ORE_FIELDS] = {0, 0, 0, 0, 0};
  unsigned dim = 0, k;
  unsigned long long r;
  qk_comp_stats st;

  CHECK(qk_model_create(&model));
  CHECK(qk_model_add_component(model, &stock, &s1));
  CHECK(qk_model_add_component(model, &process, &p1));
  CHECK(qk_model_add_component(model, &stock, &s2));
  CHECK(qk_model_set_vector_n(model, s1, IRON_ORE_FIELDS, IRON_ORE_TOTAL_MASK, 10.0, 100.0, init1));
  CHECK(qk_model_set_vector_n(model, p1, IRON_ORE_FIELDS, IRON_ORE_TOTAL_MASK, 0.0, 0.0, 0));
  CHECK(qk_model_set_vector_n(model, s2, IRON_ORE_FIELDS, IRON_ORE_TOTAL_MASK, 10.0, 100.0, init2));
  CHECK(qk_model_set_dist(model, p1, 0, &time10, &d_time));
  CHECK(qk_model_set_dist(model, p1, 1, &quantity, &d_qty));
  CHECK(qk_model_connect(model, s1, p1, -1));
  CHECK(qk_model_connect(model, p1, s2, -1));

  cfg.n_reps = n_reps;
  cfg.rep_offset = 0;
  cfg.base_seed = 123456789ull;
  cfg.device = 0;
  cfg.log_capacity = 0;
  cfg.threads_per_block = 0;
  cfg.flags = 0;
  cfg.stream = 0;
  cfg.lanes_per_replication = 0;
  cfg.reserved = 0;
  CHECK(qk_batch_create(model, &cfg, &batch));
  CHECK(qk_batch_init(batch, 0));
  CHECK(qk_batch_step_until(batch, 300ull * 1000000000ull));
  CHECK(qk_batch_synchronize(batch));

  /* replication 0 in full, then the mean content of the second stockpile over a sample of replications */
  CHECK(qk_batch_read_vector_n(batch, 0, s2, ore, QK_VEC_MAX_DIM, &dim));
  CHECK(qk_batch_comp_stats(batch, 0, 1, p1, &st));
  printf("replication 0: %llu batches moved, MyStock2 = {fe %.17g, other_elements %.17g, magnetite %.17g, hematite %.17g, "
         "limonite %.17g} (%u fields), fe%% %.6f\n",
         (unsigned long long)st.count, ore[FE], ore[OTHER_ELEMENTS], ore[MAGNETITE], ore[HEMATITE], ore[LIMONITE], dim,
         100.0 * ore[FE] / (ore[FE] + ore[OTHER_ELEMENTS]));
  for (r = 0; r < n_reps && r < 64; ++r) {
    CHECK(qk_batch_read_vector_n(batch, r, s2, ore, QK_VEC_MAX_DIM, &dim));
    for (k = 0; k < IRON_ORE_FIELDS; ++k) sum[k] += ore[k];
  }
  printf("mean of %llu replications: MyStock2 total %.6f\n", r, (sum[FE] + sum[OTHER_ELEMENTS]) / (double)r);
  qk_batch_destroy(batch);
  qk_model_destroy(model);
  return 0;
}
