"""Prometheus text exposition of the device-side series (N3): what `/metrics` on :9110 serves for the
series of pkg/gateway/metrics/metrics.go that are functions of the request stream. Label values come from the same
objects the reference's collector is called with (qos.Namespace, qos.User, model): one qos entry == one label set.
Durations (gateway_request_duration_seconds, gateway_response_process_duration_milliseconds) are wall-clock
observations and stay with the host."""
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
