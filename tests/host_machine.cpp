// host_machine.cpp — TEST-ONLY build of arks_b200/csrc/json_machine.cuh with g++ so the device state
// machines can be fuzzed against the oracle on CPU (tests/test_machine_vs_oracle.py). Not shipped.
#include <stddef.h>
#include <string.h>

#include "../arks_b200/csrc/json_engine.cuh"

using namespace arks;

// same bulk loop the kernels run, with units read straight from memory (zero padded past the end)
template <class M>
static void feed(M& m, const uint8_t* body, size_t len) {
  uint32_t pos = 0;
  // exercise window boundaries like the tiled kernels do: consume in 128-byte windows
  for (uint32_t wbeg = 0; wbeg < len; wbeg += 128) {
    uint32_t lim = (uint32_t)(len < wbeg + 128 ? len : wbeg + 128);
    consume_t(m, pos, lim, [&](uint32_t u) {
      Unit16 q;
      uint8_t tmp[16] = {0};
      size_t o = (size_t)u * 16;
      size_t n = o < len ? (len - o < 16 ? len - o : 16) : 0;
      memcpy(tmp, body + o, n);
      for (int k = 0; k < 16; k++) tmp[k] = k < (int)n ? tmp[k] : (uint8_t)(0xA5 ^ k);  // garbage past the end
      memcpy(q.w, tmp, 16);
      return q;
    });
    if (m.dead()) break;
  }
}

extern "C" {
int hm_parse_request(const uint8_t* body, size_t len, uint8_t* model_out, size_t cap, size_t* model_len, int* stream,
                     int* so_present, int* iu) {
  static thread_local JsonT m;
  static thread_local uint32_t stk[kStackWords];
  static thread_local JsonCold cold;
  m.init(K_REQ, body, stk, &cold, host_json_tables());
  feed(m, body, len);
  *stream = cold.stream3;
  *so_present = cold.so_present;
  *iu = cold.iu3;
  size_t k = 0;
  if (m.ok_at_end()) {
    if (cold.m_esc)
      decode_span(body + cold.m_start, cold.m_rawlen, [&](uint8_t b) {
        if (k < cap) model_out[k] = b;
        k++;
      });
    else {
      k = cold.m_rawlen;
      memcpy(model_out, body + cold.m_start, k < cap ? k : cap);
    }
  }
  *model_len = k;
  return m.ok_at_end() ? 0 : 1;
}
int hm_parse_response(const uint8_t* body, size_t len, size_t* model_nonempty, int64_t usage[3]) {
  static thread_local JsonT m;
  static thread_local uint32_t stk[kStackWords];
  static thread_local JsonCold cold;
  m.init(K_RESP, body, stk, &cold, host_json_tables());
  feed(m, body, len);
  *model_nonempty = cold.m_rawlen > 0;
  for (int k = 0; k < 3; k++) usage[k] = cold.usage[k];
  return m.ok_at_end() ? 0 : 1;
}
int hm_parse_sse(const uint8_t* body, size_t len, int64_t usage[3]) {
  static thread_local SseT m;
  static thread_local uint32_t stk[kStackWords];
  static thread_local JsonCold cold;
  m.init(body, stk, &cold, host_json_tables());
  feed(m, body, len);
  bool ok = m.finish((uint32_t)len);
  for (int k = 0; k < 3; k++) usage[k] = m.usage[k];
  return ok ? 0 : 1;
}
}
