"""Prometheus text exposition of the device-side series (N3): what `/metrics` on :9110 serves for the
series of pkg/gateway/metrics/metrics.go that are functions of the request stream. Label values come from the same
objects the reference's collector is called with (qos.Namespace, qos.User, model): one qos entry == one label set.
Durations (gateway_request_duration_seconds, gateway_response_process_duration_milliseconds) are wall-clock
observations and stay with the host: `HostMetrics` below, fed by the ext_proc server where the reference calls its
collector (gateway.go:118,129; handle_response.go:100-106)."""
import threading

import numpy as np

from . import abi

RULE_NAMES = ("rpm", "rpd", "tpm", "tpd")           # ratelimiter rule names, pkg/gateway/ratelimiter/types.go:35-47
BUCKETS = [str(1 << k) for k in range(17)] + ["+Inf"]  # prometheus.ExponentialBuckets(1, 2, 17), metrics.go:66


def _esc(v: str) -> str:
    return v.replace("\\", "\\\\").replace("\n", "\\n").replace('"', '\\"')


def exposition(tables, rows: np.ndarray) -> str:
    """rows: Gateway.snapshot_metrics(). Series that were never touched are not emitted (a CounterVec child only exists
    after its first WithLabelValues)."""
    out = []

    def labels(q, **extra):
        tok = tables.qos_token[q]
        kv = {"namespace": tables.token_namespace[tok], "user": tables.token_user[tok], "model": tables.qos_model_name[q]}
        kv.update(extra)
        return "{" + ",".join(f'{k}="{_esc(str(v))}"' for k, v in kv.items()) + "}"

    out.append("# TYPE gateway_requests_total counter")
    for q in np.flatnonzero(rows[:, abi.METRIC_MESSAGES]):
        out.append(f"gateway_requests_total{labels(q, status='200')} {int(rows[q, abi.METRIC_MESSAGES])}")
    out.append("# TYPE gateway_rate_limit_hits_total counter")
    for q, r in zip(*np.nonzero(rows[:, abi.METRIC_HITS:abi.METRIC_HITS + 4])):
        out.append(f"gateway_rate_limit_hits_total{labels(q, rule_type=RULE_NAMES[r])} {int(rows[q, abi.METRIC_HITS + r])}")
    recorded = rows[:, abi.METRIC_HIST_IN:abi.METRIC_HIST_IN + abi.METRIC_HIST_BUCKETS].sum(axis=1)
    out.append("# TYPE gateway_token_usage counter")
    for q in np.flatnonzero(recorded):
        # the reference labels these two vectors (namespace, user, token, type) but passes the model as `token`
        tok = tables.qos_token[q]
        base = {"namespace": tables.token_namespace[tok], "user": tables.token_user[tok], "token": tables.qos_model_name[q]}
        for k, ty in enumerate(("input", "output")):
            lab = "{" + ",".join(f'{a}="{_esc(str(b))}"' for a, b in {**base, "type": ty}.items()) + "}"
            out.append(f"gateway_token_usage{lab} {int(rows[q, abi.METRIC_USAGE + k])}")
    out.append("# TYPE gateway_token_distribution histogram")
    for q in np.flatnonzero(recorded):
        tok = tables.qos_token[q]
        base = {"namespace": tables.token_namespace[tok], "user": tables.token_user[tok], "token": tables.qos_model_name[q]}
        for k, (ty, col) in enumerate((("input", abi.METRIC_HIST_IN), ("output", abi.METRIC_HIST_OUT))):
            cum = np.cumsum(rows[q, col:col + abi.METRIC_HIST_BUCKETS])
            for le, c in zip(BUCKETS, cum):
                lab = "{" + ",".join(f'{a}="{_esc(str(b))}"' for a, b in {**base, "type": ty, "le": le}.items()) + "}"
                out.append(f"gateway_token_distribution_bucket{lab} {int(c)}")
            lab = "{" + ",".join(f'{a}="{_esc(str(b))}"' for a, b in {**base, "type": ty}.items()) + "}"
            out.append(f"gateway_token_distribution_sum{lab} {int(rows[q, abi.METRIC_USAGE + k])}")
            out.append(f"gateway_token_distribution_count{lab} {int(cum[-1])}")
    return "\n".join(out) + "\n"


def _go_float(x: float) -> str:
    """strconv.FormatFloat(x, 'g', -1, 64) for the values that occur here (bucket bounds and sums)"""
    if x == int(x) and abs(x) < 1e21:
        return str(int(x))
    return repr(float(x))


class HostMetrics:
    """The series of metrics.go that depend on the wall clock or on non-200 statuses, kept by the host:

      gateway_request_duration_seconds{namespace,user,model}              RecordRequest, collector.go:35-38
      gateway_requests_total{...,status} for status != "200"              (the "200" children are counted on the device, one
                                                                          per response-body message: ARKS_METRIC_MESSAGES)
      gateway_response_process_duration_milliseconds{namespace,user,model} RecordRespProcessingTime, collector.go:47-49

    RecordRequest runs once per response-headers message with :status 500 and once per response-body message
    (gateway.go:118,129) with `time.Since(requestStart)` truncated to whole milliseconds, in seconds."""
    REQ_BUCKETS = (0.1, 0.5, 1, 2, 5, 10, 20, 30, 45, 60)              # metrics.go:42
    RESP_MS_BUCKETS = (1, 5, 10, 50, 100, 200, 500, 1000, 2000, 5000)  # metrics.go:52

    def __init__(self):
        self._mu = threading.Lock()
        self.requests = {}   # (ns, user, model, status) -> count, status != "200"
        self.req_hist = {}   # (ns, user, model) -> [bucket counts..., +Inf, sum]
        self.resp_hist = {}

    @staticmethod
    def _observe(table, key, buckets, v):
        h = table.setdefault(key, [0] * (len(buckets) + 1) + [0.0])
        for i, b in enumerate(buckets):
            if v <= b:
                h[i] += 1
                break
        else:
            h[len(buckets)] += 1
        h[-1] += v

    def record_request(self, namespace, user, model, elapsed_s: float, status: int):
        seconds = int(elapsed_s * 1000) / 1000  # float64(time.Since(start).Milliseconds()) / 1000
        with self._mu:
            if status != 200:
                k = (namespace, user, model, str(status))
                self.requests[k] = self.requests.get(k, 0) + 1
            self._observe(self.req_hist, (namespace, user, model), self.REQ_BUCKETS, seconds)

    def record_resp_processing(self, namespace, user, model, elapsed_s: float):
        with self._mu:
            self._observe(self.resp_hist, (namespace, user, model), self.RESP_MS_BUCKETS, float(int(elapsed_s * 1000)))

    def exposition(self) -> str:
        out = []
        lab = lambda names, vals, **extra: "{" + ",".join(
            f'{k}="{_esc(str(v))}"' for k, v in list(zip(names, vals)) + list(extra.items())) + "}"
        with self._mu:
            for k, c in sorted(self.requests.items()):
                out.append(f"gateway_requests_total{lab(('namespace', 'user', 'model', 'status'), k)} {c}")
            for name, table, buckets in (("gateway_request_duration_seconds", self.req_hist, self.REQ_BUCKETS),
                                         ("gateway_response_process_duration_milliseconds", self.resp_hist, self.RESP_MS_BUCKETS)):
                if table:
                    out.append(f"# TYPE {name} histogram")
                for k, h in sorted(table.items()):
                    cum = 0
                    for b, c in zip(list(buckets) + ["+Inf"], h[:-1]):
                        cum += c
                        le = b if b == "+Inf" else _go_float(b)
                        out.append(f"{name}_bucket{lab(('namespace', 'user', 'model'), k, le=le)} {cum}")
                    out.append(f"{name}_sum{lab(('namespace', 'user', 'model'), k)} {_go_float(h[-1])}")
                    out.append(f"{name}_count{lab(('namespace', 'user', 'model'), k)} {cum}")
        return "\n".join(out) + ("\n" if out else "")
