/* ork_internal.h — oracle internals. TEST INFRASTRUCTURE ONLY (see arks_oracle.h). */
#ifndef ORK_INTERNAL_H
#define ORK_INTERNAL_H
#include <stddef.h>
#include <stdint.h>

#include "arks_oracle.h"

typedef struct {
  uint8_t* buf;
  size_t len, cap; /* len may exceed cap: bytes beyond cap are counted, not stored */
  int hash_on;     /* readFieldHash mode: hash instead of store */
  uint64_t h;
} ork_sink;

int ork_json_request(const uint8_t* body, size_t len, ork_sink* model, int* stream, int* so_present,
                     int* include_usage, uint32_t span[2]);
int ork_json_response(const uint8_t* body, size_t len, ork_sink* model, int64_t usage[3]);
int ork_sse_chunk(const uint8_t* body, size_t len, int64_t usage[3]);
void ork_usage_from_value(const uint8_t* b, size_t i, size_t e, int64_t usage[3]);

#endif
