/*
 * arks_oracle.h — CPU oracle for the arks gateway ext_proc hot path.  TEST INFRASTRUCTURE ONLY.
 *
 * This is a plain-C restatement of the reference's (scitix/arks @ e1732a9) Go decision logic with the
 * Redis counters replaced by in-memory int64s. Only tests/, __graft_entry__.smoke() and bench.py's
 * cpu_baseline / --impl reference legs may load it. The product library (arks_b200/csrc) never does.
 *
 * PARITY STATUS: the reference cannot be compiled here (no Go toolchain) and its own
 * tests pin no decisions. The oracle is pinned against every fixture the reference
 * does hold (tests/test_oracle_golden.py); everything that lives in un-vendored third-party modules
 * (json-iterator v1.1.12, openai-go v0.1.0-beta.3 ssestream/apijson, gjson v1.14.4, encoding/json) is
 * restated from their published algorithms: for those parts **parity is unpinned** (DESIGN.md §4).
 */
#ifndef ARKS_ORACLE_H
#define ARKS_ORACLE_H
#include "../include/arks_gateway.h"
#ifdef __cplusplus
extern "C" {
#endif

typedef struct ork ork;

ork* ork_create(const arks_tables* t);
void ork_destroy(ork* o);
int ork_reload(ork* o, const arks_tables* t);
int ork_update_endpoint_weights(ork* o, uint32_t endpoint, uint32_t n, const int32_t* w);

/* serial batch application: request i is fully applied (check, then increment) before i+1 */
int ork_request_batch(ork* o, const arks_request_batch* b, arks_request_result* r);
int ork_response_batch(ork* o, const arks_response_batch* b, arks_response_result* r);
/* N4 (arks_set_precharge): the mode, and the prompt token counts of the NEXT request batch (the oracle has no tokenizer) */
void ork_set_precharge(ork* o, int on);
void ork_set_estimates(ork* o, const uint32_t* est, uint32_t n);
/* tenant-sharded multi-thread variants (cpu baseline): shard = hash(namespace) % nthreads */
int ork_request_batch_mt(ork* o, const arks_request_batch* b, arks_request_result* r, int nthreads);
int ork_response_batch_mt(ork* o, const arks_response_batch* b, arks_response_result* r, int nthreads);

int ork_snapshot_quota(ork* o, int64_t* usage);
int ork_set_quota_usage(ork* o, uint32_t quota, const int64_t usage[3]);
int ork_incr_quota_usage(ork* o, uint32_t quota, const int64_t delta[3]);
int ork_snapshot_rate(ork* o, int64_t now_unix, int64_t* counters);
/* syncQuotaUsage for one ArksQuota (arks_impl.go:226-296); restore 0 = the reference (zeroes on outdated), 1 = repaired */
int ork_sync_quota_usage(ork* o, uint32_t quota, uint32_t* status_present, int64_t status_used[3], int restore);
/* the Prometheus series of include/arks_gateway.h (ARKS_METRIC_*), always accumulated: n_qos * ARKS_METRIC_COLS */
int ork_snapshot_metrics(ork* o, int64_t* rows);

/* ---- unit-level entry points (tests) ---- */
/* HandleRequestBody's JSON decode (handle_request.go:87-104): returns 0 ok, 1 error.
 * model_out receives up to model_cap decoded bytes; *model_len the full decoded length.
 * tri-state outs: 0 = nil, 1 = false, 2 = true. so_present: stream_options pointer non-nil. */
int ork_parse_request_body(const uint8_t* body, size_t len, uint8_t* model_out, size_t model_cap,
                           size_t* model_len, int* stream, int* so_present, int* include_usage);
/* non-stream response decode (handle_response.go:156-182): 0 ok, 1 unmarshal error.
 * *model_len = decoded model length. usage[3] = prompt, completion, total. */
int ork_parse_response_body(const uint8_t* body, size_t len, size_t* model_len, int64_t usage[3]);
/* one SSE chunk (handle_response.go:113-133): 0 ok, 1 stream error. usage[3] as above (zeros if none) */
int ork_parse_sse_chunk(const uint8_t* body, size_t len, int64_t usage[3]);
/* the events eventStreamDecoder.Next dispatches for one chunk, serialised as {u32 type_len, u32 data_len,
 * u32 n_data_lines, type, data} records (data carries the '\n' the decoder appends per data line). Returns the event
 * count, -1 on a scanner error, -2 if `out` is too small. Pinned against openai-python's SSEDecoder (same Stainless
 * decoder family as openai-go's ssestream) in tests/test_decoder_pins.py. */
int ork_sse_events(const uint8_t* body, size_t len, uint8_t* out, size_t cap, size_t* used);
/* getWindowStart, ratelimiter/cache_key.go:73-80 */
int64_t ork_window_start(int64_t now_unix, int rule);
/* CacheKeyGenerator.Generate, ratelimiter/cache_key.go:42-71 and quota/cache_key.go:40-58. Return length. */
size_t ork_rate_key(const char* prefix, const char* ns, const char* user, const char* model, int rule,
                    int64_t now_unix, char* out, size_t cap);
size_t ork_quota_key(const char* prefix, const char* ns, const char* quota, int type, char* out, size_t cap);
/* Envoy weighted pick restated (A12): cumulative walk of r mod sum(w); -1 if sum == 0 */
int32_t ork_weighted_pick(const int32_t* w, uint32_t n, uint64_t r);
/* HandleRequestHeaders bearer extraction (handle_request.go:38-46) */
size_t ork_extract_bearer(const uint8_t* const* keys, const size_t* key_lens, const uint8_t* const* values,
                          const size_t* value_lens, size_t n_headers, const uint8_t** token);

#ifdef __cplusplus
}
#endif
#endif
