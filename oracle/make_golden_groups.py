"""Golden rank lists for inferix_b200.parallel_state.rank_groups: output of the reference's own RankGenerator
(/root/reference/inferix/distributed/parallel_state.py:193-234) for a sweep of (tp, cp, pp, dp, order).  Test
infrastructure; run in the build container (the reference tree is not available on the GPU box):

    python oracle/make_golden_groups.py        # writes tests/golden/parallel_groups.json
"""
import importlib.util
import itertools
import json
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
spec = importlib.util.spec_from_file_location("ref_parallel_state", "/root/reference/inferix/distributed/parallel_state.py")
ref = importlib.util.module_from_spec(spec)
spec.loader.exec_module(ref)

KINDS = ["dp", "dp-cp", "cp", "tp-pp", "tp", "tp-cp", "pp", "tp-cp-dp", "tp-dp"]
cases = []
for order in ("tp-cp-pp-dp", "tp-pp-dp-cp", "tp-dp-pp-cp", "tp-cp-dp-pp"):
    for tp, cp, pp, dp in itertools.product((1, 2), (1, 2, 4, 8), (1, 2), (1, 2, 3)):
        if tp * cp * pp * dp > 48:
            continue
        gen = ref.RankGenerator(tp=tp, dp=dp, pp=pp, cp=cp, order=order)
        cases.append({"sizes": {"tp": tp, "cp": cp, "pp": pp, "dp": dp}, "order": order,
                      "groups": {k: gen.get_ranks(k) for k in KINDS}})
# an order that omits size-1 axes (the generator appends them), as dist_init's callers may pass
gen = ref.RankGenerator(tp=1, dp=2, pp=1, cp=4, order="cp-dp")
cases.append({"sizes": {"tp": 1, "cp": 4, "pp": 1, "dp": 2}, "order": "cp-dp",
              "groups": {k: gen.get_ranks(k) for k in KINDS}})
out = ROOT / "tests" / "golden" / "parallel_groups.json"
out.write_text(json.dumps({"source": "reference RankGenerator.get_ranks", "cases": cases}, separators=(",", ":")))
print(f"{len(cases)} cases -> {out} ({out.stat().st_size} bytes)")
