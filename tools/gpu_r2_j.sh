#!/bin/bash
# round-2 call J (1 GPU): PCIe ceiling probe, smoke(), the whole GPU suite, the default bench line + launch list
set -u
mkdir -p gpurun_out
timeout 120 tools/_build/pcie_probe > gpurun_out/j_pcie_probe.jsonl 2>&1; cat gpurun_out/j_pcie_probe.jsonl
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/j_pytest.log 2>&1
echo "pytest rc=$?"; tail -4 gpurun_out/j_pytest.log
timeout 600 python bench.py > gpurun_out/j_bench_n1.json 2> gpurun_out/j_bench_n1.err
echo "bench rc=$?"; python - <<'PY'
import json
d = json.load(open('gpurun_out/j_bench_n1.json'))
print(d['value'], d['kernel_ms'], d['roofline']['frac'], d['e2e']['value'], d['e2e_pageable']['value'], d['cpu_baseline'].get('value'), d['clocks'])
for k, v in d['configs'].items(): print(k, v.get('ms'), v.get('frac'), v.get('parity_ok'), v.get('error'))
PY
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 2>/dev/null | cut -c1-300
