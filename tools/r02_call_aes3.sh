#!/bin/bash
set -u
out=gpurun_out/r02aes7
mkdir -p "$out"
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_fuzz.py tests/test_chstone_aes.py -m gpu -x -q -k "aes or host_call or fuzz or reference_entry" > "$out/pytest.log" 2>&1; echo "pytest rc=$?"; tail -3 "$out/pytest.log"
{
for nc in 1 2 3; do
echo "== nc$nc inj0"; python tools/profile_target.py --kernel aes --nc $nc --log2n 24 --iters 10 --time
echo "== nc$nc inj1 threshold 0"; python tools/profile_target.py --kernel aes --nc $nc --log2n 24 --iters 10 --time --threshold 0
echo "== nc$nc inj1 2^-10"; python tools/profile_target.py --kernel aes --nc $nc --log2n 24 --iters 10 --time --inject 0.0009765625
done
echo "== nc2 inj1 2^-7"; python tools/profile_target.py --kernel aes --nc 2 --log2n 24 --iters 10 --time --inject 0.0078125
echo "== dec nc2 inj0"; python tools/profile_target.py --kernel aes --nc 2 --log2n 24 --iters 10 --time --aes-mode 1
echo "== dec nc2 inj1 2^-10"; python tools/profile_target.py --kernel aes --nc 2 --log2n 24 --iters 10 --time --aes-mode 1 --inject 0.0009765625
echo "== dec nc3 inj0"; python tools/profile_target.py --kernel aes --nc 3 --log2n 24 --iters 10 --time --aes-mode 1
echo "== enck nc2"; python tools/profile_target.py --kernel aes --nc 2 --log2n 24 --iters 10 --time --aes-mode 2
echo "== deck nc2"; python tools/profile_target.py --kernel aes --nc 2 --log2n 24 --iters 10 --time --aes-mode 3
} > "$out/timings.txt" 2>&1
grep -E "^==|best" "$out/timings.txt" | cut -c1-110
timeout 600 ncu --set full --clock-control none -k regex:xmr_aes128_enc_nc2_inj1 -c 1 -o "$out/aes_nc2_inj1" python tools/profile_target.py --kernel aes --nc 2 --log2n 24 --iters 2 --inject 0.0009765625 > "$out/ncu_aes.log" 2>&1
python tools/ncu_summary.py "$out/aes_nc2_inj1.ncu-rep" "$out/aes_nc2_inj1.json" > /dev/null 2>&1; rm -f "$out/aes_nc2_inj1.ncu-rep"
python -c "
import json
d=json.load(open('$out/aes_nc2_inj1.json'))[0];print({k:d.get(k) for k in ('kernel','duration','pipe_alu_pct','pipe_lsu_pct','warp_insts','issue_per_cycle_per_smsp','registers','dram_read','dram_write')})"
