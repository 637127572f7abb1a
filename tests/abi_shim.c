/* abi_shim.c — TEST-ONLY stand-in for libarksgw.so on machines without a GPU: the three entry points the C++ host
 * (host/cpp/arks_host.cc) calls, answered by the CPU oracle. It lets the batcher's concurrency, ordering and the
 * ext_proc state machine be tested here; the GPU tests run the same host code against the real library. */
#include <stdlib.h>

#include "../include/arks_gateway.h"
#include "../oracle/arks_oracle.h"

#include <string.h>

#define SHIM_SLOTS 4
struct slot_res {
  uint32_t n;
  uint8_t *reason, *detail, *flags, *counted;
  int32_t *qos, *token, *pick;
  int64_t *cur, *lim, *usage;
  uint32_t *moff, *mlen, *bpe;
  int rc;
};
struct arks_ctx {
  ork* o;
  int cur;
  uint32_t generation;
  int precharge;
  struct slot_res rq[SHIM_SLOTS], rs[SHIM_SLOTS];
};

int arks_shim_create(const arks_tables* t, arks_ctx** out) {
  arks_ctx* c = (arks_ctx*)calloc(1, sizeof *c);
  c->o = ork_create(t);
  if (!c->o) { free(c); return ARKS_E_BAD_TABLE; }
  c->generation = 1;
  *out = c;
  return 0;
}
uint32_t arks_table_generation(const arks_ctx* c) { return c->generation; }
/* the shim keeps no key history: rows of an older generation are answered ARKS_R_QOS_GONE (the product re-maps them) */
int arks_load_tables(arks_ctx* c, const arks_tables* t) {
  int rc = ork_reload(c->o, t);
  if (!rc) c->generation++;
  return rc;
}
/* prepare/commit: the shim has no device image to build ahead; the tables pointer must stay valid until the commit (the
 * host library calls both inside one LoadTables) */
struct arks_prepared { const arks_tables* t; uint32_t base; };
int arks_prepare_tables(arks_ctx* c, const arks_tables* t, arks_prepared** out) {
  arks_prepared* p = (arks_prepared*)malloc(sizeof *p);
  p->t = t;
  p->base = c->generation;
  *out = p;
  return 0;
}
int arks_commit_tables(arks_ctx* c, arks_prepared* p) {
  if (p->base != c->generation) return ARKS_E_INVALID_ARG;
  int rc = arks_load_tables(c, p->t);
  if (!rc) free(p);
  return rc;
}
void arks_discard_prepared(arks_ctx* c, arks_prepared* p) { (void)c; free(p); }
/* the object store lives in the product library only (its host half is tested through tests/host_machine.cpp) */
int arks_config_prepare(arks_ctx* c, arks_prepared** out) { (void)c; (void)out; return ARKS_E_INVALID_ARG; }
void* arks_shim_oracle(arks_ctx* c) { return c->o; } /* for snapshots in tests */
void arks_shim_destroy(arks_ctx* c) {
  if (!c) return;
  ork_destroy(c->o);
  free(c);
}
/* N4 through the shim: the oracle has no tokenizer, so the stand-in "prompt count" of a request is body_len / 4 — what matters
 * to the host tests is that the estimate the request phase reports comes back with the stream's response */
int arks_set_precharge(arks_ctx* c, int on) { c->precharge = on != 0; ork_set_precharge(c->o, on); return 0; }
static int shim_request(arks_ctx* c, const arks_request_batch* b, arks_request_result* r) {
  uint32_t* est = NULL;
  if (c->precharge && b->n) {
    est = (uint32_t*)malloc(4 * (size_t)b->n);
    for (uint32_t i = 0; i < b->n; i++) est[i] = b->body_len[i] / 4;
    ork_set_estimates(c->o, est, b->n);
  }
  int rc = ork_request_batch(c->o, b, r);
  if (r->bpe_count) for (uint32_t i = 0; i < b->n; i++) r->bpe_count[i] = est ? est[i] : 0;
  free(est);
  return rc;
}
int arks_submit_request_batch(arks_ctx* c, const arks_request_batch* b, arks_request_result* r) { return shim_request(c, b, r); }
int arks_submit_response_batch(arks_ctx* c, const arks_response_batch* b, arks_response_result* r) { return ork_response_batch(c->o, b, r); }
void* arks_alloc_pinned(size_t bytes) { return aligned_alloc(64, (bytes + 127) & ~(size_t)63); }
void arks_free_pinned(void* p) { free(p); }
size_t arks_extract_bearer(const uint8_t* const* keys, const size_t* key_lens, const uint8_t* const* values,
                           const size_t* value_lens, size_t n_headers, const uint8_t** token) {
  return ork_extract_bearer(keys, key_lens, values, value_lens, n_headers, token);
}

/* asynchronous surface: the oracle answers at submit time, the answer is parked per slot until arks_wait_* */
int arks_select_slot(arks_ctx* c, int slot) {
  if (slot < 0 || slot >= SHIM_SLOTS) return ARKS_E_INVALID_ARG;
  c->cur = slot;
  return 0;
}
static void grow(struct slot_res* s, uint32_t n) {
  if (s->n >= n && s->reason) return;
  s->n = n;
  s->reason = realloc(s->reason, n); s->detail = realloc(s->detail, n); s->flags = realloc(s->flags, n); s->counted = realloc(s->counted, n);
  s->qos = realloc(s->qos, 4 * (size_t)n); s->token = realloc(s->token, 4 * (size_t)n); s->pick = realloc(s->pick, 4 * (size_t)n);
  s->cur = realloc(s->cur, 8 * (size_t)n); s->lim = realloc(s->lim, 8 * (size_t)n); s->usage = realloc(s->usage, 24 * (size_t)n);
  s->moff = realloc(s->moff, 4 * (size_t)n); s->mlen = realloc(s->mlen, 4 * (size_t)n); s->bpe = realloc(s->bpe, 4 * (size_t)n);
}
int arks_submit_request_async(arks_ctx* c, const arks_request_batch* b) {
  struct slot_res* s = &c->rq[c->cur];
  grow(s, b->n ? b->n : 1);
  s->n = b->n;
  arks_request_result r = {s->reason, s->detail, s->flags, s->qos, s->token, s->pick, s->cur, s->lim, s->moff, s->mlen, s->bpe};
  s->rc = shim_request(c, b, &r);
  return s->rc;
}
int arks_wait_request(arks_ctx* c, int slot, arks_request_result* out) {
  struct slot_res* s = &c->rq[slot];
  size_t n = s->n;
  memcpy(out->reason, s->reason, n); memcpy(out->detail, s->detail, n); memcpy(out->flags, s->flags, n);
  memcpy(out->qos, s->qos, 4 * n); memcpy(out->token, s->token, 4 * n); memcpy(out->pick, s->pick, 4 * n);
  memcpy(out->cur_usage, s->cur, 8 * n); memcpy(out->limit_max, s->lim, 8 * n);
  if (out->model_off) memcpy(out->model_off, s->moff, 4 * n);
  if (out->model_len) memcpy(out->model_len, s->mlen, 4 * n);
  if (out->bpe_count) memcpy(out->bpe_count, s->bpe, 4 * n);
  return s->rc;
}
int arks_submit_response_async(arks_ctx* c, const arks_response_batch* b) {
  struct slot_res* s = &c->rs[c->cur];
  grow(s, b->n ? b->n : 1);
  s->n = b->n;
  arks_response_result r = {s->reason, s->counted, s->usage};
  if (b->gen) { /* rows of an older generation: their qos index means nothing in these tables */
    int32_t* q = (int32_t*)malloc(4 * (size_t)(b->n ? b->n : 1));
    for (uint32_t i = 0; i < b->n; i++) q[i] = b->gen[i] == c->generation ? b->qos[i] : -1;
    arks_response_batch b2 = *b;
    b2.qos = q;
    s->rc = ork_response_batch(c->o, &b2, &r);
    free(q);
  } else {
    s->rc = ork_response_batch(c->o, b, &r);
  }
  return s->rc;
}
int arks_wait_response(arks_ctx* c, int slot, arks_response_result* out) {
  struct slot_res* s = &c->rs[slot];
  size_t n = s->n;
  memcpy(out->reason, s->reason, n); memcpy(out->counted, s->counted, n); memcpy(out->usage, s->usage, 24 * n);
  return s->rc;
}
