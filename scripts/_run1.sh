set -u
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --tb=short -p no:cacheprovider -x > gpurun_out/pytest.log 2>&1; echo "pytest exit $?"; tail -n 25 gpurun_out/pytest.log
for cfg in "3 32 1" "3 64 1" "4 32 1" "3 32 0" "3 32 2" "6 32 2" "8 32 2" "3 32 3"; do
  set -- $cfg
  JSLP_DEBUG=1 NO_FARM=1 NO_CPU=1 KNAP_NODES=200 SPEC=32 REPS=1 NODE_SLOTS=$1 SLOT_STEPS=$2 STEP_VARIANT=$3 timeout 300 python scripts/mip_bench.py > gpurun_out/mip_s$1_$2_v$3.log 2> gpurun_out/mip_s$1_$2_v$3.err; echo "slots $cfg exit $?"
  python - <<PY
import json,re
rec=json.loads(open("gpurun_out/mip_s$1_$2_v$3.log").read().strip().splitlines()[-1])
print({k:rec[k] for k in ("node_phase_ms","node_phase_node_lps_per_s","host_root_ms","rounds","node_lps","launches")})
ts=[(int(m.group(1)),float(m.group(2))) for m in re.finditer(r"running (\d+) loads \d+: (\d+) us", open("gpurun_out/mip_s$1_$2_v$3.err").read())]
import collections
d=collections.defaultdict(list)
for r,t in ts: d[r].append(t)
print({r:(len(v), round(sum(v)/len(v))) for r,v in sorted(d.items())})
PY
done
