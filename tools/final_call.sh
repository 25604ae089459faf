#!/bin/bash
# One GPU call: A/B of the search kernel's memory-path variants (tools/stage_ab.py), the contract bench with the plain kernel and
# with the best variant, then the whole GPU test suite and the default bench line with whichever of the two won.
# Everything lands in gpurun_out/r2d/.  Time guards keep the call under its limit.
O=gpurun_out/r2d
mkdir -p $O
LIMIT=${LIMIT:-450}
left() { echo $(( LIMIT - SECONDS )); }
timeout 100 python tools/stage_ab.py --rounds 4 > $O/stage_ab.json 2> $O/stage_ab.err
echo "stage_ab rc=$? t=$SECONDS" | tee $O/final_call.log
BEST=$(python - <<'PY'
import json
try:
    d = json.loads(open("gpurun_out/r2d/stage_ab.json").read().strip().splitlines()[-1])
    print(d["best"] if d["best_over_plain"] < 0.985 else 0)
except Exception:
    print(0)
PY
)
echo "best variant: $BEST" | tee -a $O/final_call.log
FQB_SEARCH_VAR=0 timeout 80 python bench.py --no-cpu-baseline --no-cli > $O/bench_var0.json 2> $O/bench_var0.err
echo "bench var0 rc=$? t=$SECONDS" | tee -a $O/final_call.log
WIN=0
if [ "$BEST" != "0" ]; then
    FQB_SEARCH_VAR=$BEST timeout 80 python bench.py --no-cpu-baseline --no-cli > $O/bench_var$BEST.json 2> $O/bench_var$BEST.err
    echo "bench var$BEST rc=$? t=$SECONDS" | tee -a $O/final_call.log
    WIN=$(python - $BEST <<'PY'
import json, sys
try:
    a = json.loads(open("gpurun_out/r2d/bench_var0.json").read().strip().splitlines()[-1])
    b = json.loads(open("gpurun_out/r2d/bench_var%s.json" % sys.argv[1]).read().strip().splitlines()[-1])
    print(1 if b["value"] > 1.01 * a["value"] and b["e2e"]["value"] > 1.0 * a["e2e"]["value"] else 0)
except Exception:
    print(0)
PY
)
fi
[ "$WIN" = "1" ] && USE=$BEST || USE=0
echo "win: $WIN  -> variant $USE for the suite and the default bench" | tee -a $O/final_call.log
export FQB_SEARCH_VAR=$USE
timeout 110 python bench.py > $O/bench_default.json 2> $O/bench_default.err
echo "bench default (with cli + cpu_baseline legs) rc=$? t=$SECONDS" | tee -a $O/final_call.log
if [ $(left) -gt 120 ]; then
    timeout $(( $(left) - 15 )) python -m pytest tests -m gpu -x -q > $O/gputests_final.txt 2>&1
    echo "suite rc=$? t=$SECONDS" | tee -a $O/final_call.log
    tail -3 $O/gputests_final.txt | tee -a $O/final_call.log
fi
if [ $(left) -gt 45 ]; then
    timeout 40 python tools/stage_ab.py --read-len 150 --pairs 131072 --rounds 3 --variants 0,1,5,7,15 > $O/stage_ab_150.json 2> $O/stage_ab_150.err
    echo "stage_ab 150 rc=$? t=$SECONDS" | tee -a $O/final_call.log
fi
if [ $(left) -gt 30 ]; then
    timeout 28 python bench.py --no-cpu-baseline --no-cli --config wgs_mix --steps 10 > $O/bench_wgs.json 2> $O/bench_wgs.err
    echo "bench wgs rc=$? t=$SECONDS" | tee -a $O/final_call.log
fi
echo "done t=$SECONDS" | tee -a $O/final_call.log
