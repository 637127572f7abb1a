"""Regenerates tests/golden/*.json from the reference tree (run in the build container only:
`python tests/golden/make_golden.py /root/reference`). The fixtures are small excerpts of the reference's
own sample objects / README bodies; /root/reference does not exist on the GPU box, so they are committed."""
import json
import os
import re
import sys

import yaml

ref = sys.argv[1] if len(sys.argv) > 1 else "/root/reference"
out = os.path.dirname(os.path.abspath(__file__))

docs = [d for d in yaml.safe_load_all(open(os.path.join(ref, "examples/quickstart/quickstart.yaml"))) if d]
pick = lambda kind: [{"metadata": {k: v for k, v in d["metadata"].items() if k in ("name", "namespace")},
                      "spec": d["spec"]} for d in docs if d["kind"] == kind]
readme = open(os.path.join(ref, "README.md")).read()
req = re.search(r"-d '(\{\"model\": \"qwen-7b\".*?)'", readme).group(1)
resp = re.search(r"Expected response\s*``` json\s*(\{.*?\n\})\s*```", readme, re.S).group(1)
tokens, quotas, endpoints = pick("ArksToken"), pick("ArksQuota"), pick("ArksEndpoint")
for e in endpoints:  # quickstart omits the namespace on the endpoint: kubectl applies it to "default"
    e["metadata"].setdefault("namespace", "default")
json.dump({"tokens": tokens, "quotas": quotas, "endpoints": endpoints, "request_body": req, "response_body": resp},
          open(os.path.join(out, "quickstart.json"), "w"), indent=1)

st = yaml.safe_load(open(os.path.join(ref, "config/samples/arks_v1_arkstoken.yaml")))
tok = {"metadata": {"name": st["metadata"]["name"], "namespace": st["metadata"].get("namespace", "default")},
       "spec": st["spec"]}
model = tok["spec"]["qos"][0]["arksEndpoint"]["name"]
sq = yaml.safe_load(open(os.path.join(ref, "config/samples/arks_v1_arksquota.yaml")))
se = yaml.safe_load(open(os.path.join(ref, "config/samples/arks_v1_arksendpoint.yaml")))
json.dump({"tokens": [tok], "model": model,
           "quotas": [{"metadata": {"name": sq["metadata"]["name"], "namespace": sq["metadata"]["namespace"]},
                       "spec": sq["spec"]}],
           "endpoints": [{"metadata": {"name": se["metadata"]["name"], "namespace": "default"},
                          "spec": {"defaultWeight": se["spec"]["defaultWeight"]}}]},
          open(os.path.join(out, "sample_token.json"), "w"), indent=1)
print("wrote quickstart.json, sample_token.json")
