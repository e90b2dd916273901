#!/bin/bash
# round 2, first GPU call: parity tests, the new bench line, topology facts, memcheck of the super-tone state rows
mkdir -p gpurun_out
nvidia-smi topo -m > gpurun_out/r02_topo.txt 2>&1
(nproc; lscpu | head -30; ls /sys/devices/system/node/ 2>/dev/null; for d in /sys/bus/pci/devices/*; do if [ -f $d/numa_node ] && grep -qi 0x10de $d/vendor 2>/dev/null; then echo $d $(cat $d/numa_node) $(cat $d/class); fi; done; free -g | head -2; cat /proc/self/status | grep -i allowed) >> gpurun_out/r02_topo.txt 2>&1
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/r02_pytest_gpu.log 2>&1; tail -5 gpurun_out/r02_pytest_gpu.log
timeout 900 python -X faulthandler bench.py --steps 10 --warmup 3 > gpurun_out/r02_bench_a.json 2> gpurun_out/r02_bench_a.err; tail -3 gpurun_out/r02_bench_a.err
python - <<'PY'
import json
try:
    d=json.load(open('gpurun_out/r02_bench_a.json'))
    print('value', d['value'], 'ms', d['ms_per_step'], 'frac', d['roofline']['frac'], 'kern', d['roofline']['kernel_ms'])
    print('e2e', d['e2e']['value'], d['e2e']['ms_per_step'], d['e2e'].get('numa'), 'g711', d['e2e'].get('g711_ulaw',{}).get('value'))
    print('parity', d['parity_check'])
    print('cpu', d['cpu_baseline'])
    for k,v in (d.get('configs') or {}).items():
        print(k, {kk: v.get(kk) for kk in ('error','value','ms_per_step','parity_check','cpu_baseline')}, 'e2e', (v.get('e2e') or {}).get('value'), 'roof', (v.get('roofline') or {}).get('kernel_ms'), (v.get('roofline') or {}).get('frac'))
except Exception as e:
    print('bench parse failed', e)
PY
timeout 600 compute-sanitizer --tool memcheck --error-exitcode 3 python -m pytest tests/test_gpu_wire.py -k "state_rows or wide" -x -q > gpurun_out/r02_memcheck_super_tone.log 2>&1; echo memcheck rc $?; tail -4 gpurun_out/r02_memcheck_super_tone.log
