# A/B of library variants: bash tools/ab.sh tag1 tag2 ...   ("default" = the in-tree library)
for t in "$@"; do
  if [ "$t" = default ]; then unset CMBPO_B200_LIB; else export CMBPO_B200_LIB=tools/lib_$t.so; fi
  timeout 200 python bench.py --steps 6 --warmup 3 --cpu-batch 0 --no-e2e > gpurun_out/ab_$t.json 2> gpurun_out/ab_$t.err || echo "bench $t failed"
  python - "$t" <<'PY'
import json, sys
t = sys.argv[1]
try:
    d = json.load(open("gpurun_out/ab_%s.json" % t))
    b = d["breakdown_ms_per_step"]
    print("%-10s value %.2fM  ms/step %.3f  K1 %.4f ms  dyn %.2f pol %.2f row %.2f  gae %.3f of HBM" % (t, d["value"] / 1e6, d["ms_per_step"], d["roofline"]["launch_ms"], b["dynamics_gemm_chain"], b["policy_pass"], b["row_kernel"], d["gae"]["frac"]))
except Exception as e:
    print(t, "no result", e)
PY
done
