"""Synthetic OpenAI-shape gateway traffic (BASELINE.json configs 2-5). Seeded and deterministic.

Used by bench.py, __graft_entry__.smoke() and the parity tests; it only *produces bytes* — decisions come from
the CUDA library (product) or from the oracle (checker).
"""
from __future__ import annotations

import numpy as np

from . import abi
from .abi import RequestBatch, RequestResult, ResponseBatch, pack_blobs, pack_concat
from .tables import Tables, simple_endpoint, simple_quota

WORDS = ("the of and to in is that for it as was with be by on not he this are or his from at which but have an had "
         "they you were their one all we can her has there been if more when will would who so no out up said what "
         "its about than into them only other time new some could these two may first then do any like my now over "
         "such our man me even most made after also did many off before must well back through years much where your "
         "way down should because long each just state people those too how little good world make very year still "
         "see own work men day get here old life both between being under never know same last another while us might "
         "great since against right came take used himself few house use place during high without again home around "
         "small however found mrs thought went say part once general upon every left war don does got united number "
         "hand course water until always away public something fact less though far put head think called set almost "
         "enough end took night government yet better told nothing eyes find going look asked later point knew city "
         "next program business give group toward days young let room").split()

MODEL = "qwen-7b"


def _text(rng: np.random.Generator, nbytes: int) -> str:
    out, n = [], 0
    while n < nbytes + 16:
        w = WORDS[int(rng.integers(len(WORDS)))]
        out.append(w)
        n += len(w) + 1
    return " ".join(out)[:nbytes]


def chat_request_body(rng, size: int = 1024, stream: bool = False) -> bytes:
    """`{"model":"qwen-7b","messages":[{"role":"user","content":"..."}]}` padded to exactly `size` bytes."""
    head = '{"model":"%s","messages":[{"role":"user","content":"' % MODEL
    tail = '"}]' + (',"stream":true,"stream_options":{"include_usage":true}' if stream else "") + "}"
    return (head + _text(rng, size - len(head) - len(tail)) + tail).encode()


# ---- varied traffic: what a gateway in front of many different clients sees ------------------------------------
_ESC = ['\\n', '\\"', '\\t', '\\u00e9', '\\u4f60\\u597d', '\\\\', '\\/']


def _prose(rng, nbytes: int) -> str:
    """text for a JSON string: words, with the escapes real prompts carry (newlines, quotes, \\u sequences) and some
    raw UTF-8, at irregular positions"""
    out, n = [], 0
    while n < nbytes:
        r = rng.random()
        if r < 0.04:
            w = _ESC[int(rng.integers(len(_ESC)))]
        elif r < 0.06:
            w = ("é", "你好", "—", "🙂")[int(rng.integers(4))]
        else:
            w = WORDS[int(rng.integers(len(WORDS)))]
        out.append(w)
        n += len(w.encode()) + 1
    return " ".join(out)


class ClientApp:
    """One application talking to the gateway: it always sends the same parameters in the same order, the same system
    prompt and the same JSON style (that is what an SDK call site produces); what changes per request is the
    conversation — how many turns and how long each one is."""

    def __init__(self, rng):
        params = []
        if rng.random() < 0.5: params.append('"temperature":%s' % ("0.7", "1", "0", "0.25")[int(rng.integers(4))])
        if rng.random() < 0.5: params.append('"max_tokens":%d' % int(rng.integers(16, 4096)))
        if rng.random() < 0.3: params.append('"top_p":0.%d' % int(rng.integers(1, 10)))
        if rng.random() < 0.2: params.append('"user":"u-%06d"' % int(rng.integers(10**6)))
        if rng.random() < 0.15: params.append('"stop":["\\n\\n","###"]')
        if rng.random() < 0.1: params.append('"presence_penalty":0,"frequency_penalty":0.5')
        if rng.random() < 0.1: params.append('"response_format":{"type":"json_object"}')
        if rng.random() < 0.1: params.append('"n":1,"seed":%d' % int(rng.integers(1 << 31)))
        self.explicit_no_stream = rng.random() < 0.3
        rng.shuffle(params)
        self.params = params
        self.cut = int(rng.integers(len(params) + 1))
        self.system = _prose(rng, int(rng.integers(40, 400))) if rng.random() < 0.5 else None
        self.sep = ("", " ")[int(rng.random() < 0.15)]       # some clients pretty-print a little
        self.model_first = rng.random() >= 0.3               # model is not always the first key
        self.max_turns = int(rng.integers(1, 7))


_APPS = {}


def client_apps(n_apps: int = 256, seed: int = 0xC11E):
    apps = _APPS.get((n_apps, seed))
    if apps is None:
        r = np.random.default_rng(seed)
        apps = _APPS[(n_apps, seed)] = [ClientApp(r) for _ in range(n_apps)]
    return apps


def chat_request_body_varied(rng, mean_size: int = 1024, stream: bool = False, app: ClientApp = None) -> bytes:
    """An OpenAI chat request from one of a few hundred client applications (ClientApp): per-app parameters, key order,
    system prompt and JSON style; per-request 1..max_turns turns of very different lengths, total size spread around
    `mean_size` (roughly 0.4x-1.7x)."""
    if app is None:
        apps = client_apps()
        app = apps[int(rng.integers(len(apps)))]
    target = int(mean_size * (0.34 + 1.32 * rng.random()))
    params = list(app.params)
    tail = []
    if stream:
        tail = ['"stream":true', '"stream_options":{"include_usage":true}']
    elif app.explicit_no_stream:
        tail = ['"stream":false']
    n_turns = int(rng.integers(1, app.max_turns + 1))
    roles = [("user", "assistant")[k % 2] for k in range(n_turns)]
    if roles[-1] != "user":
        roles.append("user")
    budget = max(target - (len(app.system) if app.system else 0) - 80, 32)
    shares = rng.random(len(roles)) ** 2 + 0.02
    shares = shares / shares.sum()
    msgs = (['{"role":"system","content":"%s"}' % app.system] if app.system else []) + \
           ['{"role":"%s","content":"%s"}' % (r, _prose(rng, max(int(budget * sh) - 40, 1))) for r, sh in zip(roles, shares)]
    sep = app.sep
    model = ['"model":"%s"' % MODEL]
    messages = ['"messages":%s[%s]' % (sep, ("," + sep).join(msgs))]
    if app.model_first:
        fields = model + params[:app.cut] + messages + params[app.cut:] + tail
    else:
        fields = params[:app.cut] + messages + model + params[app.cut:] + tail
    return ("{" + ("," + sep).join(fields) + "}").encode()


def chat_response_body_varied(rng, prompt: int, completion: int, mean_size: int = 600) -> bytes:
    """A non-streaming completion as vLLM / SGLang / OpenAI-compatible servers write it: field sets and order differ,
    content length spread around `mean_size`."""
    content = _prose(rng, max(int(mean_size * (0.3 + 1.4 * rng.random())) - 395, 4))
    style = int(rng.integers(3))
    usage_fields = ['"prompt_tokens":%d' % prompt, '"completion_tokens":%d' % completion, '"total_tokens":%d' % (prompt + completion)]
    if style == 0:
        usage_fields = [usage_fields[0], usage_fields[2], usage_fields[1], '"prompt_tokens_details":null']
    elif style == 1:
        usage_fields.append('"prompt_tokens_details":{"cached_tokens":%d},"completion_tokens_details":{"reasoning_tokens":0}'
                            % int(rng.integers(0, prompt + 1)))
    usage = '"usage":{%s}' % ",".join(usage_fields)
    msg = '"message":{"role":"assistant","content":"%s"%s}' % (
        content, ("", ',"reasoning_content":null,"tool_calls":null', ',"refusal":null,"annotations":[]')[style])
    choice = '{"index":0,%s,"logprobs":null,"finish_reason":"%s"%s}' % (
        msg, ("stop", "length")[int(rng.random() < 0.2)], (',"matched_stop":151645', "", ',"stop_reason":null')[style])
    head = ['"id":"chatcmpl-%08x"' % int(rng.integers(1 << 32)), '"object":"chat.completion"', '"created":17%08d' % int(rng.integers(10**8)),
            '"model":"%s"' % MODEL]
    if style == 2:
        head.append('"system_fingerprint":"fp_%06x","service_tier":"default"' % int(rng.integers(1 << 24)))
    body = head + ['"choices":[%s]' % choice, usage]
    if style == 1:  # usage before choices
        body = head + [usage, '"choices":[%s]' % choice]
    return ("{" + ",".join(body) + "}").encode()


_FILL = {}


def _filler(rng, nbytes: int) -> str:
    """Cheap filler text: one of 64 cached strings per length (generation speed for 64k-response waves)."""
    pool = _FILL.get(nbytes)
    if pool is None:
        pr = np.random.default_rng(nbytes)
        pool = _FILL[nbytes] = [_text(pr, nbytes) for _ in range(64)]
    return pool[int(rng.integers(64))]


def chat_response_body(rng, prompt: int, completion: int, size: int = 600) -> bytes:
    head = ('{"id":"chatcmpl-%08x","object":"chat.completion","created":1700000000,"model":"%s","choices":[{"index":0,'
            '"message":{"role":"assistant","content":"' % (int(rng.integers(1 << 32)), MODEL))
    tail = ('","reasoning_content":null,"tool_calls":null},"logprobs":null,"finish_reason":"stop","matched_stop":151645}],'
            '"usage":{"prompt_tokens":%d,"total_tokens":%d,"completion_tokens":%d,"prompt_tokens_details":null}}'
            % (prompt, prompt + completion, completion))
    return (head + _filler(rng, max(size - len(head) - len(tail), 8)) + tail).encode()


def sse_response_chunks(rng, prompt: int, completion: int, total: int = 4096, n_chunks: int = 4):
    """A chat.completion.chunk stream of ~`total` bytes cut on frame boundaries into `n_chunks` chunks; the last
    chunk carries the usage frame (choices == []) and `data: [DONE]`."""
    cid = "chatcmpl-%08x" % int(rng.integers(1 << 32))
    frame = lambda s: ('data: {"id":"%s","object":"chat.completion.chunk","created":1700000000,"model":"%s",'
                       '"choices":[{"index":0,"delta":{"content":"%s"},"logprobs":null,"finish_reason":null}],'
                       '"usage":null}\n\n' % (cid, MODEL, s)).encode()
    last = ('data: {"id":"%s","object":"chat.completion.chunk","created":1700000000,"model":"%s","choices":[],'
            '"usage":{"prompt_tokens":%d,"total_tokens":%d,"completion_tokens":%d}}\n\ndata: [DONE]\n\n'
            % (cid, MODEL, prompt, prompt + completion, completion)).encode()
    per = total // n_chunks
    chunks = []
    for c in range(n_chunks):
        budget = per - (len(last) if c == n_chunks - 1 else 0)
        parts, used = [], 0
        while True:
            f = frame(_text(rng, int(rng.integers(4, 40))))
            if used + len(f) > budget:
                break
            parts.append(f)
            used += len(f)
        if c == n_chunks - 1:
            parts.append(last)
        chunks.append(b"".join(parts))
    return chunks


class Workload:
    """N tenants = N namespaces, each with one ArksToken (one qos entry for qwen-7b, rpm/rpd/tpm/tpd limits),
    one ArksQuota (prompt/response/total) and one ArksEndpoint — SURVEY.md §8d config 2."""

    def __init__(self, n_tenants: int = 10_000, seed: int = 0xA2C5, zipf_alpha: float = 0.0, deny_target: float = 0.05,
                 n_backends: int = 3):
        rng = np.random.default_rng(seed)
        self.n_tenants = n_tenants
        self.seed = seed
        tokens, quotas, endpoints = [], [], []
        self.token_strings = []
        # limits drawn so that ~5 % of requests hit rpm and ~1 % hit the quota in a multi-wave run
        rpm = rng.integers(3, 400, n_tenants)
        tight = rng.random(n_tenants) < deny_target * 2
        rpm = np.where(tight, rng.integers(1, 6, n_tenants), rpm)
        quota_total = np.where(rng.random(n_tenants) < 0.02, rng.integers(200, 4000, n_tenants),
                               rng.integers(10**6, 10**9, n_tenants))
        for t in range(n_tenants):
            ns = "tenant-%05d" % t
            tok = "sk-%016x" % int(rng.integers(1 << 62))
            self.token_strings.append(tok.encode())
            tokens.append({"metadata": {"name": "user-%05d" % t, "namespace": ns},
                           "spec": {"token": tok, "qos": [{
                               "arksEndpoint": {"name": MODEL},
                               "rateLimits": [{"type": "rpm", "value": int(rpm[t])},
                                              {"type": "tpm", "value": int(rpm[t]) * 2000},
                                              {"type": "rpd", "value": int(rpm[t]) * 200},
                                              {"type": "tpd", "value": int(rpm[t]) * 400000}],
                               "quota": {"name": "quota-%05d" % t}}]}})
            quotas.append(simple_quota("quota-%05d" % t, ns, [("prompt", int(quota_total[t])),
                                                             ("response", int(quota_total[t]) * 5),
                                                             ("total", int(quota_total[t]) * 6)]))
            endpoints.append(simple_endpoint(MODEL, ns, default_weight=5,
                                             routes=[("model-service-%d" % k, int(rng.integers(1, 100)))
                                                     for k in range(n_backends)]))
        self.objects = (tokens, quotas, endpoints)  # the CRD-shaped objects (merged shards, incremental config tests)
        self.tables = Tables(tokens, quotas, endpoints)
        if zipf_alpha > 0:
            p = 1.0 / np.power(np.arange(1, n_tenants + 1, dtype=np.float64), zipf_alpha)
            self.popularity = p / p.sum()
        else:
            self.popularity = None
        self._tok_fixed = np.frombuffer(b"".join(self.token_strings), np.uint8).reshape(n_tenants, -1)

    def draw_tenants(self, rng, n):
        if self.popularity is None:
            return rng.integers(0, self.n_tenants, n)
        return rng.choice(self.n_tenants, size=n, p=self.popularity)

    def request_batch(self, n: int, now_unix: int, seed: int = 1, body_size: int = 1024, stream_frac: float = 0.0,
                      noise_frac: float = 0.0, n_templates: int = 0, varied: bool = False) -> RequestBatch:
        """n requests from tenants drawn by popularity. `noise_frac` mixes in malformed / unauthorised requests
        (parity tests); `n_templates` > 0 reuses that many distinct bodies (faster generation for big waves);
        `varied`: bodies differ in structure and size (mean `body_size`) instead of one fixed shape of exactly
        `body_size` bytes."""
        rng = np.random.default_rng([self.seed, seed])
        tenants = self.draw_tenants(rng, n)
        bodies, tokens = [], []
        templ = None
        make = chat_request_body_varied if varied else chat_request_body
        if n_templates:
            templ = [make(rng, body_size, stream=(k < n_templates * stream_frac)) for k in range(n_templates)]
        noise = rng.random(n) < noise_frac if noise_frac > 0 else np.zeros(n, bool)
        for i in range(n):
            tok = self.token_strings[int(tenants[i])]
            if templ is not None:
                body = templ[int(rng.integers(n_templates))]
            else:
                body = make(rng, body_size, stream=bool(rng.random() < stream_frac))
            if noise[i]:
                k = int(rng.integers(8))
                if k == 0:
                    tok = b"sk-unknown-%d" % i
                elif k == 1:
                    body = body[: int(rng.integers(1, len(body)))]
                elif k == 2:
                    body = body.replace(b'"model":"qwen-7b"', b'"model":"other-model"')
                elif k == 3:
                    body = body.replace(b'"model":"qwen-7b",', b"")
                elif k == 4:
                    body = body.replace(b'"messages"', b'"stream":true,"messages"', 1)
                elif k == 5:
                    body = body[:-1] + b',"MODEL":"qw\\u0065n-7b","stream":null}'
                elif k == 6:
                    body = body + b"  \n\x00trailing"
                else:
                    body = body.replace(b'"role":"user"', b'"role":"user","n":-01,"x":[1,2.5e3,{"k":null}]', 1)
            bodies.append(body)
            tokens.append(tok)
        pick = rng.integers(0, 1 << 63, n, dtype=np.uint64)
        return RequestBatch.from_lists(bodies, tokens, now_unix, pick_rand=pick)

    def response_batch(self, req_result: RequestResult, now_unix: int, seed: int = 2, body_size: int = 600,
                       sse_total: int = 4096, noise_frac: float = 0.0, varied: bool = False, n_templates: int = 0) -> ResponseBatch:
        """One response (non-stream: the complete body; stream: its final SSE chunk preceded by its content chunks)
        for every admitted request of `req_result`. `varied`: completions differ in field set, order and size (mean
        `body_size`); `n_templates` > 0 (with `varied`) draws every non-stream body from that many pre-generated ones."""
        rng = np.random.default_rng([self.seed, seed])
        bodies, qos, flags = [], [], []
        templ = None
        if varied and n_templates:
            templ = [chat_response_body_varied(rng, int(rng.integers(50, 401)), int(rng.integers(1, 513)), body_size)
                     for _ in range(n_templates)]
        for i in np.nonzero(req_result.reason == abi.R_OK)[0]:
            prompt, completion = int(rng.integers(50, 401)), int(rng.integers(1, 513))
            if req_result.flags[i] & 1:
                for ch in sse_response_chunks(rng, prompt, completion, sse_total):
                    bodies.append(ch)
                    qos.append(int(req_result.qos[i]))
                    flags.append(abi.RESP_STREAM)
            else:
                if templ is not None:
                    body = templ[int(rng.integers(n_templates))]
                else:
                    body = (chat_response_body_varied if varied else chat_response_body)(rng, prompt, completion, body_size)
                if noise_frac and rng.random() < noise_frac:
                    k = int(rng.integers(4))
                    if k == 0:
                        body = body[: int(rng.integers(1, len(body)))]
                    elif k == 1:
                        body = body.replace(b'"model":"qwen-7b",', b"")
                    elif k == 2:
                        body = body.replace(b'"total_tokens":', b'"total_tokens":0,"x":')
                    else:
                        body = body.replace(b'"usage":{', b'"usage":{"total_tokens":"7",')
                bodies.append(body)
                qos.append(int(req_result.qos[i]))
                flags.append(abi.RESP_END_OF_STREAM)
        return ResponseBatch.from_lists(bodies, qos, flags, now_unix)
