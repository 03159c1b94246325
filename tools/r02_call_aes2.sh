#!/bin/bash
set -u
out=gpurun_out/r02aes2
mkdir -p "$out"
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_fuzz.py -m gpu -x -q -k "aes or host_call or fuzz or reference_entry or smoke" > "$out/pytest.log" 2>&1; echo "pytest rc=$?"; tail -3 "$out/pytest.log"
timeout 300 python -m pytest tests/test_board_b200_flow.py tests/test_campaign.py -m gpu -x -q > "$out/pytest2.log" 2>&1; echo "pytest2 rc=$?"; tail -3 "$out/pytest2.log"
{
for nc in 1 2 3; do
echo "== nc$nc inj0"; python tools/profile_target.py --kernel aes --nc $nc --log2n 24 --iters 10 --time
echo "== nc$nc inj1 threshold 0"; python tools/profile_target.py --kernel aes --nc $nc --log2n 24 --iters 10 --time --threshold 0
echo "== nc$nc inj1 2^-10"; python tools/profile_target.py --kernel aes --nc $nc --log2n 24 --iters 10 --time --inject 0.0009765625
done
echo "== nc2 inj1 2^-7"; python tools/profile_target.py --kernel aes --nc 2 --log2n 24 --iters 10 --time --inject 0.0078125
echo "== dec nc2 inj0"; python tools/profile_target.py --kernel aes --nc 2 --log2n 24 --iters 10 --time --aes-mode 1
echo "== dec nc2 inj1 2^-10"; python tools/profile_target.py --kernel aes --nc 2 --log2n 24 --iters 10 --time --aes-mode 1 --inject 0.0009765625
echo "== enck nc2"; python tools/profile_target.py --kernel aes --nc 2 --log2n 24 --iters 10 --time --aes-mode 2
echo "== nc2 table all-zero"; python tools/profile_target.py --kernel aes --nc 2 --log2n 24 --iters 10 --time --table-density 0
} > "$out/ablate.txt" 2>&1
grep -E "^==|best" "$out/ablate.txt" | cut -c1-110
timeout 600 ncu --set full --clock-control none -k regex:xmr_aes128_enc_nc2_inj1 -c 1 -o "$out/aes_nc2_inj1" python tools/profile_target.py --kernel aes --nc 2 --log2n 24 --iters 2 --inject 0.0009765625 > "$out/ncu_aes.log" 2>&1
python tools/ncu_summary.py "$out/aes_nc2_inj1.ncu-rep" "$out/aes_nc2_inj1.json" > /dev/null 2>&1; rm -f "$out/aes_nc2_inj1.ncu-rep"
timeout 600 ncu --set full --clock-control none -k regex:xmr_aes128_enc_nc2_inj0 -c 1 -o "$out/aes_nc2_inj0" python tools/profile_target.py --kernel aes --nc 2 --log2n 24 --iters 2 > "$out/ncu_aes0.log" 2>&1
python tools/ncu_summary.py "$out/aes_nc2_inj0.ncu-rep" "$out/aes_nc2_inj0.json" > /dev/null 2>&1; rm -f "$out/aes_nc2_inj0.ncu-rep"
python -c "
import json
for f in ('aes_nc2_inj0','aes_nc2_inj1'):
    d=json.load(open('$out/'+f+'.json'))[0];print({k:d.get(k) for k in ('kernel','duration','pipe_alu_pct','pipe_lsu_pct','warp_insts','issue_per_cycle_per_smsp','registers','dram_read','dram_write')})"
