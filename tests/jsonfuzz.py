"""Seeded generators of (mostly) JSON / SSE byte strings for differential tests: well-formed documents in the
OpenAI chat shape, plus structure-aware and byte-level mutations that exercise every grammar quirk the
oracle restates (duplicate keys, case-folded keys, escapes in keys, null keys, odd numbers, control bytes,
truncation, trailing bytes, deep nesting)."""
import random

WS = [b"", b"", b"", b" ", b"\n", b"\t", b"\r", b"  "]
KEYS_REQ = [b"model", b"stream", b"stream_options", b"include_usage", b"messages", b"MODEL", b"Stream",
            b"STREAM_OPTIONS", b"Include_Usage", b"mod\\u0065l", b"str\\u0065am", b"x", b"temperature", b"", b"usage",
            b"stream_option", b"m\\u006Fdel", b"include\\u005fusage"]
KEYS_RESP = [b"model", b"usage", b"prompt_tokens", b"completion_tokens", b"total_tokens", b"choices", b"error",
             b"id", b"Usage", b"MODEL", b"total\\u005ftokens", b"Total_tokens", b"prompt_tokens_details", b"us\\u0061ge",
             b"ch\\u006fices", b"err\\u006fr", b"object"]
NUMS = [b"0", b"1", b"-1", b"25", b"45", b"20", b"123456789", b"9223372036854775807", b"-9223372036854775808",
        b"18446744073709551615", b"99999999999999999999", b"1.5", b"-0", b"0.0", b"1e3", b"1E+2", b"12.5e-1", b"-01",
        b"01", b"-.5", b"1.", b".5", b"+1", b"1e", b"1e+", b"--1", b"1.2.3", b"0x10", b"1e5e5", b"00", b"-", b"1-2",
        b"45.0", b"4.5e1", b"100e-2", b"7e0", b"0e5", b"0.5e1", b"3E2", b"2.50", b"1e19", b"123456789012345678"]
STRS = [b'""', b'"a"', b'"qwen-7b"', b'"hello world"', b'"\\n"', b'"\\u0041"', b'"\\ud83d\\ude00"', b'"\\ud800"',
        b'"\\ud800\\n"', b'"\\udc00\\ud800\\udc00"', b'"a\\"b"', b'"\\\\"', b'"\\/"', b'"\x01"', b'"\\x"', b'"\\u12G4"',
        b'"\\u12"', b'"tab\there"', b'"a\\\\\x02b"', b'"\xc3\xa9"', b'"\xff\xfe"', b'"12"', b'"-7"', b'"1a"', b'"-"',
        b'"Qwen-7B"', b'"\\u0071wen-7b"', b'"unterminated']
LITS = [b"null", b"true", b"false", b"nul", b"tru", b"fals", b"nulll", b"True", b"NULL", b"truee"]


class Gen:
    def __init__(self, seed):
        self.r = random.Random(seed)

    def ws(self):
        return self.r.choice(WS)

    def scalar(self):
        k = self.r.random()
        if k < 0.35:
            return self.r.choice(NUMS)
        if k < 0.7:
            return self.r.choice(STRS)
        return self.r.choice(LITS)

    def value(self, depth=0):
        k = self.r.random()
        if depth > 3 or k < 0.55:
            return self.scalar()
        if k < 0.8:
            n = self.r.randint(0, 3)
            sep = b"," if self.r.random() < 0.93 else self.r.choice([b"", b",,", b";"])
            body = sep.join(self.ws() + self.value(depth + 1) + self.ws() for _ in range(n))
            if self.r.random() < 0.04:
                body += b","
            close = b"]" if self.r.random() < 0.96 else self.r.choice([b"}", b""])
            return b"[" + self.ws() + body + close
        return self.obj(depth + 1, self.r.choice([KEYS_REQ, KEYS_RESP]))

    def key(self, keys):
        if self.r.random() < 0.03:
            return self.r.choice([b"null", b"1", b"'a'", b"true"])
        return b'"' + self.r.choice(keys) + b'"'

    def obj(self, depth, keys, forced=()):
        n = self.r.randint(0, 4)
        members = [self.ws() + self.key(keys) + self.ws() + (b":" if self.r.random() < 0.97 else b"") + self.ws() +
                   self.value(depth) + self.ws() for _ in range(n)]
        members += list(forced)
        self.r.shuffle(members)
        body = (b"," if self.r.random() < 0.95 else b"").join(members)
        if self.r.random() < 0.03:
            body += b","
        close = b"}" if self.r.random() < 0.96 else self.r.choice([b"]", b""])
        return b"{" + body + close

    def usage_obj(self):
        f = []
        for k in (b"prompt_tokens", b"completion_tokens", b"total_tokens"):
            if self.r.random() < 0.85:
                kk = k if self.r.random() < 0.9 else self.r.choice(KEYS_RESP)
                v = self.r.choice(NUMS[:12] + [b"25", b"20", b"45", b'"12"', b"true", b"false", b"null", b"{}", b"[1]"]) \
                    if self.r.random() < 0.5 else str(self.r.randint(0, 5000)).encode()
                f.append(b'"' + kk + b'":' + self.ws() + v)
        if self.r.random() < 0.2:
            f.append(b'"prompt_tokens_details":' + self.r.choice([b"null", b'{"cached_tokens":3}']))
        if self.r.random() < 0.15:
            f.append(b'"total_tokens":' + self.r.choice([b"null", b"7", b'"x"']))
        self.r.shuffle(f)
        return b"{" + b",".join(f) + b"}"

    def request(self):
        forced = []
        if self.r.random() < 0.9:
            forced.append(b'"' + self.r.choice([b"model", b"model", b"Model", b"mod\\u0065l"]) + b'":' + self.ws() +
                          self.r.choice([b'"qwen-7b"', b'"qwen-7b"', b'"m1"', b'""', b"null", b"7", b'"a\\u002db"',
                                         b'"\\ud83d\\ude00"', b'"x\\ty"']))
        if self.r.random() < 0.5:
            forced.append(b'"stream":' + self.ws() + self.r.choice([b"true", b"false", b"null", b"1", b'"true"']))
        if self.r.random() < 0.5:
            inner = self.r.choice([b'{"include_usage":true}', b'{"include_usage":false}', b"{}", b"null",
                                   b'{"include_usage":null}', b'{"x":1,"include_usage":true}', b"[]", b"3",
                                   b'{"INCLUDE_USAGE":true}', b'{"include_usage":true,"include_usage":null}'])
            forced.append(b'"stream_options":' + self.ws() + inner)
        if self.r.random() < 0.7:
            forced.append(b'"messages":[{"role":"user","content":' + self.r.choice(STRS[:12]) + b"}]")
        doc = self.obj(1, KEYS_REQ, forced)
        return self.tail(doc)

    def response(self):
        forced = []
        if self.r.random() < 0.9:
            forced.append(b'"model":' + self.r.choice([b'"qwen-7b"', b'""', b"null", b'"\\u0041"', b"5"]))
        if self.r.random() < 0.85:
            forced.append(b'"usage":' + self.ws() + (self.usage_obj() if self.r.random() < 0.85 else self.value(2)))
        if self.r.random() < 0.15:
            forced.append(b'"usage":' + self.usage_obj())
        doc = self.obj(1, KEYS_RESP, forced)
        return self.tail(doc)

    def tail(self, doc):
        k = self.r.random()
        if k < 0.8:
            return self.ws() + doc + self.ws()
        if k < 0.85:
            return doc + self.r.choice([b"x", b"{}", b"\x00", b"\x00junk", b" \x00", b","])
        if k < 0.92:
            return doc[:self.r.randint(0, len(doc))]
        return self.mutate(doc)

    def mutate(self, doc):
        b = bytearray(doc)
        for _ in range(self.r.randint(1, 3)):
            if not b:
                break
            i = self.r.randrange(len(b))
            op = self.r.random()
            if op < 0.4:
                b[i] = self.r.choice(b'{}[]",:\\ntf0-1.eE \n\x00\x1f')
            elif op < 0.7:
                del b[i]
            else:
                b.insert(i, self.r.choice(b'{}[]",:\\ntf0-1.eE \n\x00\x1f'))
        return bytes(b)

    def event_json(self):
        forced = []
        k = self.r.random()
        if k < 0.5:
            forced.append(b'"choices":' + self.r.choice([b'[{"delta":{"content":"hi"}}]', b"[]", b"null", b"{}", b"[1]",
                                                         b"[ ]", b'"x"']))
        if self.r.random() < 0.5:
            forced.append(b'"usage":' + (self.usage_obj() if self.r.random() < 0.8 else self.value(2)))
        if self.r.random() < 0.05:
            forced.append(b'"error":' + self.r.choice([b"null", b'{"message":"x"}', b"1"]))
        if self.r.random() < 0.1:
            forced.append(b'"choices":[]')
        doc = self.obj(1, KEYS_RESP, forced) if self.r.random() < 0.9 else self.value(0)
        if self.r.random() < 0.1:
            doc = self.mutate(doc)
        return doc

    def sse_chunk(self):
        out = []
        for _ in range(self.r.randint(0, 5)):
            k = self.r.random()
            nl = b"\n" if self.r.random() < 0.85 else b"\r\n"
            if k < 0.7:
                d = self.event_json()
                if self.r.random() < 0.05 and b"\n" not in d and len(d) > 4:
                    h = self.r.randint(1, len(d) - 1)
                    out.append(b"data:" + self.r.choice([b" ", b""]) + d[:h] + nl + b"data: " + d[h:] + nl)
                else:
                    out.append(self.r.choice([b"data: ", b"data:", b"data:  "]) + d.replace(b"\n", b" ") + nl)
            elif k < 0.78:
                out.append(b"data: [DONE]" + nl)
            elif k < 0.84:
                out.append(b": keep-alive" + nl)
                continue
            elif k < 0.9:
                out.append(self.r.choice([b"event: thread.run", b"event: message", b"event:thread.x", b"id: 7",
                                          b"retry: 3", b"data", b"event", b"garbage line"]) + nl)
                continue
            else:
                out.append(nl)
                continue
            if self.r.random() < 0.9:
                out.append(nl)
        body = b"".join(out)
        if self.r.random() < 0.1:
            body = body[:self.r.randint(0, len(body))]
        if self.r.random() < 0.05:
            body = self.mutate(body)
        return body
