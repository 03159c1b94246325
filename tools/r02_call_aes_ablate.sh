#!/bin/bash
# where does the AES injector time go?  (DWC, 2^24 blocks)
set -u
out=gpurun_out/r02aes
mkdir -p "$out"
{
echo "== inj0"; python tools/profile_target.py --kernel aes --nc 2 --log2n 24 --iters 10 --time
echo "== inj1 bernoulli threshold 0 (Philox evaluated, never hits)"; python tools/profile_target.py --kernel aes --nc 2 --log2n 24 --iters 10 --time --threshold 0
echo "== inj1 bernoulli 2^-10"; python tools/profile_target.py --kernel aes --nc 2 --log2n 24 --iters 10 --time --inject 0.0009765625
echo "== inj1 bernoulli 2^-13"; python tools/profile_target.py --kernel aes --nc 2 --log2n 24 --iters 10 --time --inject 0.0001220703125
echo "== inj1 bernoulli 2^-7"; python tools/profile_target.py --kernel aes --nc 2 --log2n 24 --iters 10 --time --inject 0.0078125
echo "== inj1 table all-zero (no Philox, no hits)"; python tools/profile_target.py --kernel aes --nc 2 --log2n 24 --iters 10 --time --table-density 0
echo "== inj1 table density 2^-10"; python tools/profile_target.py --kernel aes --nc 2 --log2n 24 --iters 10 --time --table-density 0.0009765625
} > "$out/ablate.txt" 2>&1
grep -E "^==|best" "$out/ablate.txt" | cut -c1-120
