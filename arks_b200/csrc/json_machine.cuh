// json_machine.cuh — per-lane byte state machines for the gateway hot path (device code, sm_100a).
//
// One lane == one body. The machine is a pushdown automaton that consumes one byte per step and never
// looks back, so the kernels can feed it from shared-memory tiles staged with cp.async while 31
// neighbouring lanes parse 31 other bodies in lock step.
//
// Three configurations of the same engine (reference call sites, paths relative to the reference tree):
//   K_REQ   request body   -> {model, stream, stream_options.include_usage}   pkg/gateway/handle_request.go:87-104
//   K_RESP  response body  -> {model, usage{prompt,completion,total}}         pkg/gateway/handle_response.go:89-93,157
//   K_EVT   one SSE event's data -> {error?, len(choices)==0, usage}          pkg/gateway/handle_response.go:113-124
// K_REQ/K_RESP follow json-iterator v1.1.12 (ConfigFastest: case-insensitive field hash, strict skip,
// last duplicate wins); K_EVT follows encoding/json validation + gjson/apijson extraction. The grammar and
// its documented divergences are the ones the oracle states (oracle/ork_json.c header).
//
// The file is host+device so that tests/ can compile it with g++ and fuzz it against the oracle on CPU
// (tests/test_machine_vs_oracle.py); the shipped library only instantiates it inside CUDA kernels.
#pragma once
#include <stdint.h>

#include "skip_dfa_tables.h"

// Table-driven validation of skipped subtrees (skip_step). Measured on B200 (round 1): it cuts executed instructions
// by 15 % but every byte then pays two DEPENDENT shared-memory lookups, and with ~3.4 resident warps per scheduler
// (64k bodies / 148 SMs) the scan kernels are latency-bound, not issue-bound: scan_response went 157 -> 212 us.
// Kept behind this switch for batches large enough to hide that latency.
#ifndef ARKS_SKIP_DFA
#define ARKS_SKIP_DFA 0
#endif

#if defined(__CUDACC__)
#define ARKS_HD __host__ __device__ __forceinline__
// rare paths are kept out of line as PURE functions of plain values (never of the machine object, whose address must
// not be taken or its fields leave the register file): keeps them from being speculated into the per-byte path
#define ARKS_OUTLINE __host__ __device__ __noinline__
#else
#define ARKS_HD inline
#define ARKS_OUTLINE inline
#endif

namespace arks {

enum : uint8_t { K_REQ = 0, K_RESP = 1, K_EVT = 2 };

// token-level states first (they share the whitespace skip), then in-token states
enum : uint8_t {
  S_TOP = 0,       // (unused: the top-level value is S_VAL with vm == VM_TOP, jsoniter readObjectStart)
  S_VAL,           // expecting a value
  S_ARR_FIRST,     // after '[': value or ']'
  S_OBJ_FIRST,     // after '{' (ReadObjectCB / encoding/json): key string or '}'
  S_OBJ_KEY,       // after ',' in such an object: key string (jsoniter also accepts the literal null)
  S_STRUCT_FIRST,  // after '{' in a jsoniter struct decoder: '}' or key
  S_STRUCT_KEY,    // after ',' in a struct decoder (readFieldHash): key
  S_COLON,         // expecting ':'
  S_AFTER,         // after a value inside a container: ',' or closer
  S_FINISH,        // top-level value complete: only whitespace may follow
  S_TOKEN_STATES,  // ---- marker
  S_STR,           // in string, no backslash seen yet
  S_STR_E,         // in string, a backslash has been seen
  S_ESC,           // byte after a backslash
  S_U,             // \uXXXX hex digits
  S_LIT,           // rest of null / true / false
  S_NUM,           // number
  S_STOP,          // jsoniter Unmarshal met a NUL byte after the value: accepted, rest ignored
  S_SKIP,          // inside a skipped subtree: the table-driven automaton of skip_dfa_tables.h runs (state in `ss`)
};

// what a string is (decides what happens at its closing quote)
enum : uint8_t { SK_VALUE_SKIP = 0, SK_VALUE_MODEL, SK_VALUE_UINT, SK_KEY_SKIP, SK_KEY_STRUCT, SK_KEY_EXACT };
// how the next value is consumed
enum : uint8_t { VM_SKIP = 0, VM_MODEL, VM_BOOL_STREAM, VM_BOOL_IU, VM_SO, VM_USAGE, VM_UINT, VM_ECHOICES, VM_TOP };
// special object one level below the top-level object
enum : uint8_t { L2_NONE = 0, L2_SO, L2_USAGE };
// RFC 8259 number DFA
enum : uint8_t { F_MINUS = 0, F_ZERO, F_INT, F_DOT, F_FRAC, F_E, F_ESIGN, F_EXP, F_DEAD };

static constexpr uint32_t kMaxDepth = 10000;  // jsoniter maxDepth == encoding/json maxNestingDepth
static constexpr uint32_t kStackWords = (kMaxDepth + 15) / 16 + 1;  // 2 bits per level (json_engine.cuh)

// jsoniter readFieldHash: int64 0x811c9dc5, ^= lower(byte), *= 0x1000193 (iter_object.go)
ARKS_HD constexpr uint64_t fhash_step(uint64_t h, uint8_t b) {
  return (h ^ (uint64_t)((b >= 'A' && b <= 'Z') ? b + 32 : b)) * 0x1000193ull;
}
ARKS_HD constexpr uint64_t fhash_lit(const char* s, int n) {
  uint64_t h = 0x811c9dc5ull;
  for (int i = 0; i < n; i++) h = fhash_step(h, (uint8_t)s[i]);
  return h;
}
static constexpr uint64_t H_MODEL = fhash_lit("model", 5);
static constexpr uint64_t H_STREAM = fhash_lit("stream", 6);
static constexpr uint64_t H_SO = fhash_lit("stream_options", 14);
static constexpr uint64_t H_IU = fhash_lit("include_usage", 13);
static constexpr uint64_t H_USAGE = fhash_lit("usage", 5);

// exact-key candidates (gjson Map(): case-sensitive, unescaped)
ARKS_HD constexpr uint64_t xhash_step(uint64_t h, uint8_t b) { return (h ^ (uint64_t)b) * 0x100000001b3ull; }
ARKS_HD constexpr uint64_t xhash_lit(const char* s, int n) {
  uint64_t h = 0xcbf29ce484222325ull;
  for (int i = 0; i < n; i++) h = xhash_step(h, (uint8_t)s[i]);
  return h;
}
// ids: 0 prompt_tokens 1 completion_tokens 2 total_tokens 3 error 4 choices 5 usage
ARKS_HD const char* xkey_str(int id) {
  switch (id) {
    case 0: return "prompt_tokens";
    case 1: return "completion_tokens";
    case 2: return "total_tokens";
    case 3: return "error";
    case 4: return "choices";
    default: return "usage";
  }
}
ARKS_HD constexpr int xkey_len(int id) { return id == 0 ? 13 : id == 1 ? 17 : id == 2 ? 12 : id == 3 ? 5 : id == 4 ? 7 : 5; }
static constexpr uint64_t X_PROMPT = xhash_lit("prompt_tokens", 13);
static constexpr uint64_t X_COMPL = xhash_lit("completion_tokens", 17);
static constexpr uint64_t X_TOTAL = xhash_lit("total_tokens", 12);
static constexpr uint64_t X_ERROR = xhash_lit("error", 5);
static constexpr uint64_t X_CHOICES = xhash_lit("choices", 7);
static constexpr uint64_t X_USAGE = xhash_lit("usage", 5);

ARKS_HD bool is_ws(uint8_t c) { return c == ' ' || c == '\n' || c == '\t' || c == '\r'; }
ARKS_HD bool is_digit(uint8_t c) { return (uint8_t)(c - '0') <= 9; }
ARKS_HD int hexval(uint8_t c) {
  if (is_digit(c)) return c - '0';
  uint8_t l = c | 0x20;
  if (l >= 'a' && l <= 'f') return l - 'a' + 10;
  return -1;
}

// ---- slow paths over an already validated raw string span (only when the span contains a backslash) ----
// Decodes jsoniter-style (readEscapedChar: surrogate pairing, lone surrogates -> U+FFFD) and feeds every
// output byte to `f`. The span is known to be well-formed.
template <class F>
ARKS_HD void decode_span(const uint8_t* p, uint32_t n, F&& f) {
  auto put_rune = [&](uint32_t r) {
    if (r <= 0x7F) {
      f((uint8_t)r);
    } else if (r <= 0x7FF) {
      f((uint8_t)(0xC0 | (r >> 6)));
      f((uint8_t)(0x80 | (r & 0x3F)));
    } else {
      if (r > 0x10FFFF || (r >= 0xD800 && r <= 0xDFFF)) r = 0xFFFD;
      if (r <= 0xFFFF) {
        f((uint8_t)(0xE0 | (r >> 12)));
        f((uint8_t)(0x80 | ((r >> 6) & 0x3F)));
        f((uint8_t)(0x80 | (r & 0x3F)));
      } else {
        f((uint8_t)(0xF0 | (r >> 18)));
        f((uint8_t)(0x80 | ((r >> 12) & 0x3F)));
        f((uint8_t)(0x80 | ((r >> 6) & 0x3F)));
        f((uint8_t)(0x80 | (r & 0x3F)));
      }
    }
  };
  auto u4 = [&](uint32_t i) {
    return (uint32_t)((hexval(p[i]) << 12) | (hexval(p[i + 1]) << 8) | (hexval(p[i + 2]) << 4) | hexval(p[i + 3]));
  };
  uint32_t i = 0;
  while (i < n) {
    uint8_t c = p[i++];
    if (c != '\\') {
      f(c);
      continue;
    }
    uint8_t e = p[i++];
    for (;;) {  // readEscapedChar, with its tail call unrolled into this loop
      if (e == 'u') {
        uint32_t r = u4(i);
        i += 4;
        if (r >= 0xD800 && r <= 0xDFFF) {
          if (i >= n || p[i] != '\\') {
            put_rune(r);
            break;
          }
          i++;  // the backslash
          e = p[i++];
          if (e != 'u') {
            put_rune(r);
            continue;  // readEscapedChar(e)
          }
          uint32_t r2 = u4(i);
          i += 4;
          if (r < 0xDC00 && r2 >= 0xDC00 && r2 < 0xE000) {
            put_rune((((r - 0xD800) << 10) | (r2 - 0xDC00)) + 0x10000);
          } else {
            put_rune(r);
            put_rune(r2);
          }
        } else {
          put_rune(r);
        }
        break;
      }
      uint8_t o = e;  // " \ / stay themselves
      if (e == 'b') o = '\b';
      else if (e == 'f') o = '\f';
      else if (e == 'n') o = '\n';
      else if (e == 'r') o = '\r';
      else if (e == 't') o = '\t';
      f(o);
      break;
    }
  }
}

// The skip automaton's tables (generated, tools/gen_skip_dfa.py). On the device the kernels copy them into shared
// memory once per block and hand the machines pointers to that copy; the host build points at the arrays directly.
struct SkipTables {
  const uint8_t* cls;     // 256: byte -> class
  const uint16_t* tab_j;  // kSkipStatesJ x kSkipClasses, jsoniter strict Skip()
  const uint16_t* tab_e;  // kSkipStatesE x kSkipClasses, encoding/json checkValid
};
#if !defined(__CUDA_ARCH__)
static const uint8_t kSkipClsHost[256] = ARKS_SKIP_CLASS_TABLE;
static const uint16_t kSkipTabJHost[kSkipStatesJ * kSkipClasses] = ARKS_SKIP_TABLE_J;
static const uint16_t kSkipTabEHost[kSkipStatesE * kSkipClasses] = ARKS_SKIP_TABLE_E;
inline SkipTables host_skip_tables() { return SkipTables{kSkipClsHost, kSkipTabJHost, kSkipTabEHost}; }
#endif

// ---- out-of-line slow paths (pure functions) ----
// exact comparison of a raw (validated) key span with candidate `id`
static ARKS_OUTLINE bool exact_verify_span(const uint8_t* p, uint32_t n, uint32_t has_esc, int id) {
  const char* s = xkey_str(id);
  int L = xkey_len(id);
  if (!has_esc) {
    if ((int)n != L) return false;
    for (int i = 0; i < L; i++)
      if (p[i] != (uint8_t)s[i]) return false;
    return true;
  }
  int k = 0;
  bool ok = true;
  decode_span(p, n, [&](uint8_t b) {
    if (k >= L || (uint8_t)s[k] != b) ok = false;
    k++;
  });
  return ok && k == L;
}
// which exact-key candidate in [lo, hi) the span equals, or -1 (khash only valid when !has_esc)
static ARKS_OUTLINE int exact_key_match(const uint8_t* p, uint32_t n, uint32_t has_esc, uint64_t khash, int lo, int hi) {
  int hit = -1;
  for (int id = lo; id < hi; id++) {
    uint64_t want = id == 0 ? X_PROMPT : id == 1 ? X_COMPL : id == 2 ? X_TOTAL : id == 3 ? X_ERROR : id == 4 ? X_CHOICES : X_USAGE;
    if ((has_esc || khash == want) && exact_verify_span(p, n, has_esc, id)) hit = id;
  }
  return hit;
}
// readFieldHash slow path: the hash of a key that contains a backslash (restarts from the first byte)
static ARKS_OUTLINE uint64_t struct_key_hash_slow(const uint8_t* p, uint32_t n) {
  uint64_t h = 0x811c9dc5ull;
  uint32_t i = 0;
  while (p[i] != '\\') h = fhash_step(h, p[i++]);
  decode_span(p + i, n - i, [&](uint8_t b) { h = fhash_step(h, b); });
  return h;
}
// gjson Result.Int for a Number token from its captured pieces
static ARKS_OUTLINE int64_t captured_number_value(uint64_t nacc, uint32_t nneg, uint32_t nplain, uint32_t novf, int32_t nfrac,
                                                   int32_t nexp, uint32_t nexpneg) {
  if (nplain) return nneg ? (int64_t)(0 - nacc) : (int64_t)nacc;  // safeInt / parseInt agree with a wrapping parse
  // mantissa * 10^(exp - frac) truncated toward zero; exact for the documented domain (<= 18 digits)
  int64_t e10 = (int64_t)(nexpneg ? -nexp : nexp) - (int64_t)nfrac;
  uint64_t m = nacc;
  bool ovf = novf;
  if (m == 0) return 0;
  while (e10 < 0 && m) { m /= 10; e10++; }
  while (e10 > 0 && !ovf) {
    if (m > 0xFFFFFFFFFFFFFFFFull / 10) ovf = true; else m *= 10;
    e10--;
  }
  if (ovf || m > 0x7FFFFFFFFFFFFFFFull) return INT64_MIN;
  return nneg ? -(int64_t)m : (int64_t)m;
}

struct JsonM {
  // ---- configuration ----
  const uint8_t* base;  // body bytes (global memory): only slow paths and key verification read it
  // NB: 32-bit fields on purpose. Byte-sized members get packed four to a register and every access then costs
  // PRMT/LOP3 shuffles (ncu: 12 % of scan_response_kernel's instructions were on `if (err | st == S_STOP)`).
  uint32_t kind;
  // ---- automaton ----
  uint32_t st, err, vm, skind, s_esc, ucnt, lit_is_key;
  uint32_t nf, tsn, tsn_any, tsn_dot, tsn_need, ncap;  // number sub-machines
  uint32_t l2, in_choices, ufield;
  uint32_t lit;      // remaining literal bytes, low byte first
  uint32_t depth;
  uint32_t sstart;   // offset of the first content byte of the current string
  uint64_t khash;
  uint32_t cur_word; // cached word of the container-type stack (bit = 1 object, 0 array)
  // number capture (usage counters; gjson Result.Int)
  uint64_t nacc;     // wrapping decimal accumulation of all digits of the mantissa
  uint32_t nneg, nplain, novf, sval_ok, sval_any, ncphase;
  int32_t nfrac, nexp;
  uint32_t nexpneg;
  // ---- outputs ----
  uint32_t m_start, m_rawlen;  // model string raw span
  uint32_t m_esc;              // raw span contains a backslash
  uint32_t stream3, so_present, iu3;
  int64_t usage[3];
  int64_t cand0, cand1, cand2;  // scalars, not an array: a dynamically indexed array would live in local memory
  uint32_t cand_set, cand_nonnull;
  uint32_t has_error_key, n_choices;
  uint32_t ss, skip_base;   // skip automaton state; depth of the subtree's root container
  const uint16_t* skt;      // skip table of this document's flavor
  const uint8_t* skc;       // byte -> class
  uint32_t* stk;  // container-type stack beyond the cached word: (kMaxDepth+31)/32+1 words, owned by the caller

  ARKS_HD void init(uint32_t k, const uint8_t* b, uint32_t* stack_words, const SkipTables& tabs) {
    base = b;
    stk = stack_words;
    skc = tabs.cls;
    skt = k == K_EVT ? tabs.tab_e : tabs.tab_j;
    ss = 0; skip_base = 0;
    kind = k;
    st = S_VAL;
    err = 0;
    vm = (k == K_EVT) ? VM_SKIP : VM_TOP;  // jsoniter's struct decoder only accepts '{' or null at the top
    skind = 0; s_esc = 0; ucnt = 0; lit_is_key = 0;
    nf = 0; tsn = 0; tsn_any = 0; tsn_dot = 0; tsn_need = 0; ncap = 0;
    l2 = L2_NONE; in_choices = 0; ufield = 255;
    lit = 0; depth = 0; sstart = 0; khash = 0; cur_word = 0;
    nacc = 0; nneg = 0; nplain = 1; novf = 0; sval_ok = 0; sval_any = 0; nfrac = 0; nexp = 0; nexpneg = 0; ncphase = 0;
    m_start = 0; m_rawlen = 0; m_esc = 0;
    stream3 = 0; so_present = 0; iu3 = 0;
    usage[0] = usage[1] = usage[2] = 0;
    cand0 = cand1 = cand2 = 0;
    cand_set = 0; cand_nonnull = 0;
    has_error_key = 0; n_choices = 0;
  }
  // restart for the next SSE event (fresh ChatCompletionChunk per event)
  ARKS_HD void reset_event() {
    const uint8_t* b = base;
    uint32_t* sw = stk;
    const SkipTables t{skc, skt, skt};
    init(K_EVT, b, sw, t);
  }

  ARKS_HD bool top_is_object() const { return (cur_word >> ((depth - 1) & 31)) & 1u; }
  ARKS_HD void push(bool is_obj) {
    if (depth >= kMaxDepth) {
      err = 1;
      return;
    }
    uint32_t nd = depth + 1;
    if (depth > 0 && ((nd - 1) >> 5) != ((depth - 1) >> 5)) {
      stk[(depth - 1) >> 5] = cur_word;
      cur_word = 0;
    }
    uint32_t bit = 1u << ((nd - 1) & 31);
    cur_word = is_obj ? (cur_word | bit) : (cur_word & ~bit);
    depth = nd;
  }
  ARKS_HD void pop() {
    uint32_t nd = depth - 1;
    if (nd > 0 && ((nd - 1) >> 5) != ((depth - 1) >> 5)) cur_word = stk[(nd - 1) >> 5];
    depth = nd;
  }
  ARKS_HD void value_done() { st = depth == 0 ? S_FINISH : S_AFTER; }

  ARKS_HD void set_cand(int64_t v) {
    if (ufield == 0) cand0 = v; else if (ufield == 1) cand1 = v; else cand2 = v;
  }
  ARKS_HD void commit_usage_cands() {
    // apijson struct decoder over node.Map(): last duplicate wins; null leaves the field untouched
    for (int f = 0; f < 3; f++)
      if ((cand_set >> f) & (cand_nonnull >> f) & 1) usage[f] = f == 0 ? cand0 : f == 1 ? cand1 : cand2;
  }
  ARKS_HD void close_container(uint8_t c) {
    bool obj = top_is_object();
    if ((c == '}') != obj) {
      err = 1;
      return;
    }
    if (depth == 2) {
      if (l2 == L2_USAGE) commit_usage_cands();
      l2 = L2_NONE;
      in_choices = 0;
    }
    pop();
    value_done();
  }

  ARKS_HD void begin_string(uint32_t kindv, uint32_t pos) {
    skind = kindv;
    s_esc = 0;
    sstart = pos + 1;
    st = S_STR;
    khash = (kindv == SK_KEY_STRUCT) ? 0x811c9dc5ull : 0xcbf29ce484222325ull;
    if (kindv == SK_VALUE_UINT) {
      nacc = 0; nneg = 0; sval_ok = 1; sval_any = 0;
    }
  }
  ARKS_HD void begin_key(uint32_t pos) {
    uint32_t k = SK_KEY_SKIP;
    if (kind != K_EVT) {
      if (depth == 1 || (depth == 2 && l2 == L2_SO)) k = SK_KEY_STRUCT;
      else if (depth == 2 && l2 == L2_USAGE) k = SK_KEY_EXACT;
    } else {
      if (depth == 1 || (depth == 2 && l2 == L2_USAGE)) k = SK_KEY_EXACT;
    }
    begin_string(k, pos);
  }
  ARKS_HD void begin_literal(uint8_t c, bool as_key) {
    // remaining bytes, low byte first
    lit = c == 'n' ? 0x006c6c75u /*ull*/ : c == 't' ? 0x00657572u /*rue*/ : 0x65736c61u /*alse*/;
    lit_is_key = as_key;
    st = S_LIT;
  }
  ARKS_HD void begin_number(uint8_t c) {
    st = S_NUM;
    tsn = (kind != K_EVT) && c != '0';  // jsoniter: '0' goes straight to ReadFloat32, others trySkipNumber first
    tsn_any = 0; tsn_dot = 0; tsn_need = 0;
    nf = c == '-' ? F_MINUS : c == '0' ? F_ZERO : F_INT;
    if (ncap) {
      nneg = c == '-';
      nacc = c == '-' ? 0 : (uint64_t)(c - '0');
      nplain = 1; novf = 0; nfrac = 0; nexp = 0; nexpneg = 0; ncphase = 0;
    }
  }
  // value start byte (whitespace already skipped). Decide first, act once: every primitive (begin_string, push, ...)
  // is inlined exactly once, which keeps the per-byte loop small enough for the instruction cache.
  ARKS_HD void begin_value(uint8_t c, uint32_t pos) {
    enum : uint32_t { W_ERR = 0, W_STR, W_LIT, W_NUM, W_OBJ, W_ARR };
    const uint32_t m = vm;
    vm = VM_SKIP;
    ncap = 0;
    // Iterator.Skip / encoding/json value
    uint32_t what = c == '"' ? W_STR : (c == 'n' || c == 't' || c == 'f') ? W_LIT : (c == '-' || is_digit(c)) ? W_NUM
                    : c == '[' ? W_ARR : c == '{' ? W_OBJ : W_ERR;
    uint32_t skv = SK_VALUE_SKIP, next = S_OBJ_FIRST, l2v = 0xffu, choices = 0;
    if (m != VM_SKIP) {
      if (m == VM_MODEL) {  // stringCodec -> ReadString: string or null
        skv = SK_VALUE_MODEL;
        if (c == 'n') { m_start = 0; m_rawlen = 0; m_esc = 0; }
        else if (c != '"') what = W_ERR;
      } else if (m == VM_BOOL_STREAM || m == VM_BOOL_IU) {  // OptionalDecoder{boolCodec}: ReadNil / ReadBool
        const uint32_t v = c == 'n' ? 0u : c == 'f' ? 1u : 2u;
        if (what != W_LIT) what = W_ERR;
        else if (m == VM_BOOL_STREAM) stream3 = v; else iu3 = v;
      } else if (m == VM_SO) {  // OptionalDecoder{oneFieldStructDecoder}
        if (c == 'n') { so_present = 0; iu3 = 0; }
        else if (c == '{') { so_present = 1; next = S_STRUCT_FIRST; l2v = L2_SO; }
        else what = W_ERR;
      } else if (m == VM_TOP) {  // readObjectStart: '{' or null
        if (c == '{') next = S_STRUCT_FIRST;
        else if (c != 'n') what = W_ERR;
      } else if (m == VM_UINT) {  // a usage counter: gjson Result.Int by JSON type
        cand_set |= 1u << ufield;
        cand_nonnull |= 1u << ufield;
        set_cand(c == 't' ? 1 : 0);
        if (c == 'n') cand_nonnull &= ~(1u << ufield);
        skv = SK_VALUE_UINT;
        ncap = 1;  // only matters for numbers
      } else if (m == VM_USAGE) {
        if (kind == K_EVT) { usage[0] = usage[1] = usage[2] = 0; }  // Map(): the last "usage" member wins outright
        cand_set = 0; cand_nonnull = 0;
        if (c == '{') l2v = L2_USAGE;
      } else {  // VM_ECHOICES
        n_choices = 0;
        choices = c == '[';
      }
    }
    if (what == W_STR) begin_string(skv, pos);
    else if (what == W_LIT) begin_literal(c, false);
    else if (what == W_NUM) begin_number(c);
    else if (what == W_ERR) err = 1;
    else {
      const bool generic = ARKS_SKIP_DFA && depth >= 1 && l2v == 0xffu && next == S_OBJ_FIRST && !choices;
      push(what == W_OBJ);
      if (generic) {  // nothing below this container is ever read: validate it with the table-driven automaton
        st = S_SKIP;
        ss = what == W_OBJ ? K_OBJ_FIRST : K_ARR_FIRST;
        skip_base = depth;
      } else if (what == W_OBJ) { if (l2v != 0xffu) l2 = l2v; st = next; }
      else { in_choices = choices ? 1u : in_choices; st = S_ARR_FIRST; }
    }
  }

  // one byte inside a skipped subtree: two table lookups; stack work only on brackets and at the end of scalars
  ARKS_HD void skip_step(uint8_t c) {
    for (;;) {
      const uint32_t t = skt[ss * kSkipClasses + skc[c]];
      const uint32_t f = t >> 6;
      ss = t & 63u;
      if (f == SKF_NONE) return;
      if (f == SKF_ERR) { err = 1; return; }
      if (f == SKF_PUSHO || f == SKF_PUSHA) { push(f == SKF_PUSHO); return; }
      if (f == SKF_DONE || f == SKF_DONE_RE) {
        ss = top_is_object() ? K_AFTER_OBJ : K_AFTER_ARR;
        if (f == SKF_DONE) return;
        continue;  // the number ended before this byte: dispatch the byte again in the after-value state
      }
      // pop
      if (top_is_object() != (f == SKF_POPO)) { err = 1; return; }
      const bool root = depth == skip_base;
      pop();
      if (root) value_done();  // back to the semantic machine
      else ss = top_is_object() ? K_AFTER_OBJ : K_AFTER_ARR;
      return;
    }
  }

  // ---- key dispatch at the closing quote ----
  ARKS_HD void struct_key_end(uint32_t pos) {
    uint64_t h = khash;
    if (s_esc) h = struct_key_hash_slow(base + sstart, pos - sstart);  // readFieldHash slow path
    if (depth == 1) {
      if (h == H_MODEL) vm = VM_MODEL;
      else if (kind == K_REQ && h == H_STREAM) vm = VM_BOOL_STREAM;
      else if (kind == K_REQ && h == H_SO) vm = VM_SO;
      else if (kind == K_RESP && h == H_USAGE) vm = VM_USAGE;
    } else {
      if (h == H_IU) vm = VM_BOOL_IU;
    }
  }
  ARKS_HD void exact_key_end(uint32_t pos) {
    const int lo = depth == 2 ? 0 : 3;
    const int hit = exact_key_match(base + sstart, pos - sstart, s_esc, khash, lo, lo + 3);
    if (hit < 0) return;
    if (hit < 3) { ufield = (uint32_t)hit; vm = VM_UINT; }
    else if (hit == 3) has_error_key = 1;
    else if (hit == 4) vm = VM_ECHOICES;
    else vm = VM_USAGE;
  }

  ARKS_HD void end_string(uint32_t pos) {
    switch (skind) {
      case SK_VALUE_SKIP: value_done(); break;
      case SK_VALUE_MODEL:
        m_start = sstart; m_rawlen = pos - sstart; m_esc = s_esc;
        value_done();
        break;
      case SK_VALUE_UINT: {  // gjson String -> parseInt(t.Str)
        int64_t v = 0;
        if (!s_esc && sval_ok && sval_any) v = nneg ? (int64_t)(0 - nacc) : (int64_t)nacc;
        set_cand(v);
        value_done();
        break;
      }
      case SK_KEY_SKIP: st = S_COLON; break;
      case SK_KEY_STRUCT: struct_key_end(pos); st = S_COLON; break;
      default: exact_key_end(pos); st = S_COLON; break;
    }
  }

  // gjson Result.Int for a Number token; called when the number ends
  ARKS_HD void finish_number_capture() {
    if (!ncap) return;
    ncap = 0;
    set_cand(captured_number_value(nacc, nneg, nplain, novf, nfrac, nexp, nexpneg));
  }

  // returns true when `c` was consumed, false when the number ended before `c` (reprocess it)
  ARKS_HD bool step_number(uint8_t c) {
    const bool numbyte = is_digit(c) || c == '.' || c == 'e' || c == 'E' || c == '+' || c == '-';
    if (tsn) {  // jsoniter trySkipNumber (iter_skip_strict.go)
      if (tsn_need) {
        if (!is_digit(c)) { err = 1; return true; }  // "missing digit after dot"
        tsn_need = 0;
      } else if (is_digit(c)) {
      } else if (c == '.') {
        if (tsn_dot) { err = 1; return true; }  // "more than one dot found in number"
        tsn_dot = 1;
        tsn_need = 1;
      } else if (c == ',' || c == ']' || c == '}' || c == ' ' || c == '\t' || c == '\n' || c == '\r') {
        if (tsn_any) {  // accepted without further validation
          finish_number_capture();
          value_done();
          return false;
        }
        tsn = 0;  // lone first char: defer to the float reader
      } else {
        tsn = 0;
      }
      tsn_any = 1;
    }
    // RFC 8259 number DFA, tracked from the first byte (divergence D1 when reached from jsoniter's fallback)
    if (!numbyte) {  // the token ends before c
      if (nf == F_ZERO || nf == F_INT || nf == F_FRAC || nf == F_EXP) {
        finish_number_capture();
        value_done();
        return false;
      }
      err = 1;
      return true;
    }
    uint32_t f = nf, nx = F_DEAD;
    if (is_digit(c)) {
      if (f == F_MINUS) nx = c == '0' ? F_ZERO : F_INT;
      else if (f == F_INT) nx = F_INT;
      else if (f == F_DOT || f == F_FRAC) nx = F_FRAC;
      else if (f == F_E || f == F_ESIGN || f == F_EXP) nx = F_EXP;
    } else if (c == '.') {
      if (f == F_ZERO || f == F_INT) nx = F_DOT;
    } else if (c == 'e' || c == 'E') {
      if (f == F_ZERO || f == F_INT || f == F_FRAC) nx = F_E;
    } else {
      if (f == F_E) nx = F_ESIGN;
    }
    nf = nx;
    if (nx == F_DEAD && !tsn) {  // readNumberAsString would swallow this byte and the parse would fail
      err = 1;
      return true;
    }
    if (ncap) {
      if (is_digit(c)) {
        if (ncphase == 2) {
          if (nexp < 100000) nexp = nexp * 10 + (c - '0');
        } else {
          if (nacc > (0xFFFFFFFFFFFFFFFFull - 9) / 10) novf = 1;
          nacc = nacc * 10 + (uint64_t)(c - '0');
          if (ncphase == 1) nfrac++;
        }
      } else {
        nplain = 0;
        if (c == '.') ncphase = 1;
        else if (c == 'e' || c == 'E') ncphase = 2;
        else if (c == '-' && ncphase == 2) nexpneg = 1;
      }
    }
    return true;
  }

  ARKS_HD void step(uint8_t c, uint32_t pos) {
    if (err | (st == S_STOP)) return;
    uint32_t s0 = st;
    if (s0 == S_SKIP) { skip_step(c); return; }
    // ---- inside a token
    if (s0 == S_STR || s0 == S_STR_E) {
      if (c == '"') { end_string(pos); return; }
      if (c == '\\') { s_esc = 1; st = S_ESC; return; }
      if (s0 == S_STR) {
        if (c < 0x20 && skind != SK_KEY_STRUCT) { err = 1; return; }  // readFieldHash has no such check
        if (skind >= SK_KEY_STRUCT) khash = skind == SK_KEY_STRUCT ? fhash_step(khash, c) : xhash_step(khash, c);
        else if (skind == SK_VALUE_UINT) {
          if (c == '-' && !sval_any && !nneg && pos == sstart) nneg = 1;
          else if (is_digit(c)) { nacc = nacc * 10 + (uint64_t)(c - '0'); sval_any = 1; }
          else sval_ok = 0;
        }
      } else if (c < 0x20 && kind == K_EVT) {
        err = 1;  // jsoniter's slow path does not check control characters, encoding/json does
      }
      return;
    }
    if (s0 > S_TOKEN_STATES) {
      if (s0 == S_LIT) {
        if (c != (uint8_t)(lit & 0xff)) { err = 1; return; }
        lit >>= 8;
        if (lit == 0) {
          if (lit_is_key) st = S_COLON; else value_done();
        }
        return;
      }
      if (s0 == S_ESC) {
        if (c == 'u') { ucnt = 4; st = S_U; }
        else if (c == '"' || c == '\\' || c == '/' || c == 'b' || c == 'f' || c == 'n' || c == 'r' || c == 't') st = S_STR_E;
        else err = 1;
        return;
      }
      if (s0 == S_U) {
        if (hexval(c) < 0) { err = 1; return; }
        if (--ucnt == 0) st = S_STR_E;
        return;
      }
      // S_NUM: when the number ends before c, c is handled below in the new (token-level) state
      if (step_number(c) || err) return;
      s0 = st;
    }
    // ---- between tokens: decide, then act once
    if (is_ws(c)) return;
    enum : uint32_t { A_ERR = 0, A_NONE, A_VALUE, A_KEY, A_CLOSE };
    uint32_t act = A_ERR;
    if (s0 == S_AFTER) {
      if (c == ',') {
        if (top_is_object()) {
          const bool strct = kind != K_EVT && (depth == 1 || (depth == 2 && l2 == L2_SO));
          st = strct ? S_STRUCT_KEY : S_OBJ_KEY;
        } else {
          st = S_VAL;
        }
        vm = VM_SKIP;
        act = A_NONE;
      } else if (c == '}' || c == ']') {
        act = A_CLOSE;
      }
    } else if (s0 == S_COLON) {
      if (c == ':') { st = S_VAL; act = A_NONE; }
    } else if (s0 == S_VAL) {
      act = A_VALUE;
    } else if (s0 == S_OBJ_KEY) {
      if (c == '"') act = A_KEY;
      else if (c == 'n' && kind != K_EVT) { begin_literal(c, true); act = A_NONE; }  // ReadString() accepts null as a key
    } else if (s0 == S_STRUCT_KEY) {
      if (c == '"') act = A_KEY;
    } else if (s0 == S_OBJ_FIRST || s0 == S_STRUCT_FIRST) {
      if (c == '"') act = A_KEY;
      else if (c == '}') act = A_CLOSE;
    } else if (s0 == S_ARR_FIRST) {
      if (c == ']') act = A_CLOSE;
      else {
        if (in_choices && depth == 2) n_choices = 1;
        act = A_VALUE;
      }
    } else {  // S_FINISH
      if (c == 0 && kind != K_EVT) { st = S_STOP; act = A_NONE; }  // frozenConfig.Unmarshal: `if c == 0` also matches a NUL byte
    }
    if (act == A_VALUE) begin_value(c, pos);
    else if (act == A_KEY) begin_key(pos);
    else if (act == A_CLOSE) close_container(c);
    else if (act == A_ERR) err = 1;
  }

  // end of input: jsoniter Unmarshal / encoding/json checkValid verdict
  ARKS_HD bool ok_at_end() const { return !err && (st == S_FINISH || st == S_STOP); }

  // ---- bulk interface used by the tiled kernels ----
  // True while the machine sits inside a string whose ordinary bytes need no per-byte action: the caller may then
  // skip ahead to the next '"', '\\' or byte < 0x20 without calling step() (skipped bytes are never control bytes,
  // so the jsoniter "control character before the first backslash" rule cannot be missed).
  ARKS_HD bool can_fast() const {
    return (st == S_SKIP) ? (ss < 4u) : ((st == S_STR_E) | ((st == S_STR) & (skind != SK_VALUE_UINT)));
  }
  // the `n` ordinary bytes being skipped are bytes [o, o+n) of the unit (q0..q3): hashed keys still need them
  ARKS_HD void skip(uint32_t n, uint32_t o, uint32_t q0, uint32_t q1, uint32_t q2, uint32_t q3) {
    if ((st == S_STR) & (skind >= SK_KEY_STRUCT)) {
      uint64_t h = khash;
      const bool fold = skind == SK_KEY_STRUCT;
      for (uint32_t k = o; k < o + n; k++) {
        const uint32_t lo = (k & 8) ? q2 : q0, hi = (k & 8) ? q3 : q1;
        const uint8_t b = (uint8_t)(((k & 4) ? hi : lo) >> (8 * (k & 3)));
        h = fold ? fhash_step(h, b) : xhash_step(h, b);
      }
      khash = h;
    }
  }
  ARKS_HD bool dead() const { return err | (st == S_STOP); }
};

// ---------------------------------------------------------------------------------------------
// SSE chunk machine: bufio.Scanner(ScanLines) + eventStreamDecoder.Next + Stream.Next
// (openai-go packages/ssestream, restated in oracle/ork_json.c: ork_sse_chunk / sse_event)
// ---------------------------------------------------------------------------------------------
struct SseM {
  JsonM ev;
  int64_t usage[3];
  uint32_t line_len;   // raw bytes of the current line (CR included)
  uint32_t name_len;   // bytes of the field name seen so far
  uint64_t name_acc;   // first 8 name bytes, little endian
  uint32_t data_pos;   // bytes of event data fed so far
  uint64_t data_head;  // first 8 bytes of the event data (the [DONE] probe)
  uint32_t ev_match;   // progress of matching the event type against "thread."
  uint32_t phase;      // 0 name, 1 just after ':', 2 value
  uint32_t field;      // 0 other, 1 data, 2 event
  uint32_t pending_cr, done, fail, thread_evt, ev_len_any;

  ARKS_HD void init(const uint8_t* base, uint32_t* stack_words, const SkipTables& tabs) {
    ev.init(K_EVT, base, stack_words, tabs);
    usage[0] = usage[1] = usage[2] = 0;
    line_len = 0; name_len = 0; name_acc = 0; data_pos = 0; data_head = 0; ev_match = 0;
    phase = 0; field = 0; pending_cr = 0; done = 0; fail = 0; thread_evt = 0; ev_len_any = 0;
  }
  ARKS_HD void feed_data(uint8_t c, uint32_t pos) {
    if (data_pos < 8) data_head |= (uint64_t)c << (8 * data_pos);
    data_pos++;
    if (!done) ev.step(c, pos);
  }
  ARKS_HD void classify_name() {
    if (name_len == 4 && (uint32_t)name_acc == 0x61746164u /*data*/) field = 1;
    else if (name_len == 5 && (name_acc & 0xFFFFFFFFFFull) == 0x746e657665ull /*event*/) field = 2;
    else field = 0;
    if (field == 2) { ev_match = 0; thread_evt = 0; }  // `event = string(value)`: the last event line wins
  }
  ARKS_HD void line_byte(uint8_t c, uint32_t pos) {
    if (phase == 0) {
      if (c == ':') { classify_name(); phase = 1; return; }
      if (name_len < 8) name_acc |= (uint64_t)c << (8 * name_len);
      name_len++;
      return;
    }
    if (phase == 1) {
      phase = 2;
      if (c == ' ') return;  // one optional space after the colon
    }
    if (field == 1) feed_data(c, pos);
    else if (field == 2) {
      static const char T[8] = {'t', 'h', 'r', 'e', 'a', 'd', '.', 0};
      if (ev_match < 7) {
        if (ev_match != 0xFFu && (uint8_t)T[ev_match] == c) { if (++ev_match == 7) thread_evt = 1; }
        else ev_match = 0xFF;
      }
    }
  }
  ARKS_HD void dispatch() {
    // Stream.Next for one event
    if (!done) {
      bool is_done = data_pos >= 6 && (data_head & 0xFFFFFFFFFFFFull) == 0x5d454e4f445bull;  // "[DONE]"
      if (is_done) done = 1;
      else if (!ev.ok_at_end() || ev.has_error_key) fail = 1;
      else {
        bool wrapped = thread_evt;
        uint32_t nc = wrapped ? 0 : ev.n_choices;
        if (nc == 0) {  // handle_response.go:119-123
          usage[0] = wrapped ? 0 : ev.usage[0];
          usage[1] = wrapped ? 0 : ev.usage[1];
          usage[2] = wrapped ? 0 : ev.usage[2];
        }
      }
    }
    ev.reset_event();
    data_pos = 0; data_head = 0; thread_evt = 0; ev_match = 0;
  }
  ARKS_HD void end_line(uint32_t pos) {
    // line content complete (CR already dropped)
    uint32_t content = line_len - (pending_cr ? 1 : 0);
    if (content == 0) {
      dispatch();
    } else {
      if (phase == 0) classify_name();  // no colon: the whole line is the field name
      if (field == 1) feed_data('\n', pos);
    }
    line_len = 0; name_len = 0; name_acc = 0; phase = 0; field = 0; pending_cr = 0;
  }
  ARKS_HD void step(uint8_t c, uint32_t pos) {
    if (fail) return;
    if (c == '\n') {
      end_line(pos);
      return;
    }
    line_len++;
    if (line_len >= 65536) {  // bufio.Scanner: token too long
      fail = 1;
      return;
    }
    if (pending_cr) {  // the held CR was not the last byte of its line
      pending_cr = 0;
      line_byte('\r', pos);
    }
    if (c == '\r') {
      pending_cr = 1;
      return;
    }
    line_byte(c, pos);
  }
  // bulk interface: inside a data value, inside a JSON string, with all line bookkeeping that depends on single
  // bytes already settled (the [DONE] probe has its 8 bytes, no CR is pending, the 64 KiB limit is far away)
  ARKS_HD bool can_fast() const {
    return (phase == 2) & (field == 1) & !done & !fail & !pending_cr & (data_pos >= 8) & (line_len < 65000) & ev.can_fast();
  }
  ARKS_HD void skip(uint32_t k, uint32_t o, uint32_t q0, uint32_t q1, uint32_t q2, uint32_t q3) {
    line_len += k;
    data_pos += k;
    ev.skip(k, o, q0, q1, q2, q3);
  }
  ARKS_HD bool dead() const { return fail; }
  // end of chunk: an unterminated last line is still a token; a pending event is dropped
  ARKS_HD bool finish(uint32_t pos) {
    if (!fail && line_len > 0) end_line(pos);
    return !fail;
  }
};

// ---------------------------------------------------------------------------------------------
// Bulk consumption: 16-byte units, SWAR search for the next byte a string cares about
// ---------------------------------------------------------------------------------------------
struct Unit16 {
  uint32_t w[4];
};

// bit 7 of byte k set  <=>  byte k of the word may be '"', '\\' or < 0x20 (other bits are garbage). The lowest flagged
// byte is always exact; higher bytes can be false positives (borrow of the subtract), which only costs a step() call on
// an ordinary byte. x ^ 0x02 maps {0x00..0x1f, '"'} onto the contiguous range 0x00..0x20, so two range tests suffice.
ARKS_HD uint32_t special_flags(uint32_t w) {
  uint32_t a = w ^ 0x02020202u, b = w ^ 0x5c5c5c5cu;
  return ((a - 0x21212121u) & ~a) | ((b - 0x01010101u) & ~b);
}
ARKS_HD uint32_t special_mask16(uint32_t q0, uint32_t q1, uint32_t q2, uint32_t q3) {
  uint32_t f0 = (special_flags(q0) >> 7) & 0x01010101u, f1 = (special_flags(q1) >> 7) & 0x01010101u;
  uint32_t f2 = (special_flags(q2) >> 7) & 0x01010101u, f3 = (special_flags(q3) >> 7) & 0x01010101u;
  return ((f0 * 0x01020408u) >> 24 & 0xfu) | ((f1 * 0x01020408u) >> 20 & 0xf0u) | ((f2 * 0x01020408u) >> 16 & 0xf00u) |
         ((f3 * 0x01020408u) >> 12 & 0xf000u);
}
ARKS_HD uint32_t first_set(uint32_t x) {
#if defined(__CUDA_ARCH__)
  return (uint32_t)__ffs((int)x) - 1u;
#else
  return (uint32_t)__builtin_ffs((int)x) - 1u;
#endif
}

// NB (measured, round 1): every loop iteration lets EVERY lane make progress — lanes inside strings swallow up to a unit,
// lanes between tokens step one byte. A variant that stepped token bytes in a tight inner loop was 40 % slower: the
// string lanes of the warp sat masked off while the token lanes looped, so the two kinds of work stopped overlapping.
// Feed bytes [pos, lim) of one body to machine `m`; `load(u)` returns 16-byte unit u of the body (bytes past the
// body's end may hold anything). Advances pos. Inside strings, units without a special byte are skipped whole.
template <class M, class L>
ARKS_HD void consume(M& m, uint32_t& pos, uint32_t lim, L&& load) {
  uint32_t cu = 0xffffffffu;
  uint32_t q0 = 0, q1 = 0, q2 = 0, q3 = 0;
  while (pos < lim) {
    const uint32_t u = pos >> 4, o = pos & 15;
    if (u != cu) {
      Unit16 q = load(u);
      q0 = q.w[0]; q1 = q.w[1]; q2 = q.w[2]; q3 = q.w[3];
      cu = u;
    }
    if (m.can_fast()) {
      const uint32_t avail = lim - pos;
      const uint32_t rest = special_mask16(q0, q1, q2, q3) >> o;
      uint32_t run = rest ? first_set(rest) : 16u - o;
      if (run > avail) run = avail;
      if (run) {
        m.skip(run, o, q0, q1, q2, q3);
        pos += run;
        // long plain stretch: keep swallowing whole clean units without re-entering the outer loop
        if (o + run == 16) {
          while (lim - pos >= 16) {
            Unit16 n = load(pos >> 4);
            const uint32_t any = (special_flags(n.w[0]) | special_flags(n.w[1]) | special_flags(n.w[2]) | special_flags(n.w[3])) & 0x80808080u;
            if (any) break;
            m.skip(16, 0, n.w[0], n.w[1], n.w[2], n.w[3]);
            pos += 16;
          }
        }
        continue;
      }
    }
    // byte o of the unit, without indexing the registers dynamically
    const uint32_t lo = (o & 8) ? q2 : q0, hi = (o & 8) ? q3 : q1;
    const uint32_t w = (o & 4) ? hi : lo;
    m.step((uint8_t)(w >> (8 * (o & 3))), pos);
    pos++;
    if (m.dead()) return;
  }
}

}  // namespace arks
