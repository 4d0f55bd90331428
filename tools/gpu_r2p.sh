#!/bin/bash
# Encrypt against the window width of Q's table; the new parity test; affine addition vs elements per inversion
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -q -x -m gpu -k "encrypt or blind or smoke or golden" > gpurun_out/r2p_pytest.log 2>&1; tail -3 gpurun_out/r2p_pytest.log
timeout 900 python tools/enc_windows.py > gpurun_out/r2p_enc_windows.json 2> gpurun_out/r2p_enc_windows.err; cat gpurun_out/r2p_enc_windows.err | cut -c1-330
for cfg in "8 37888" "64 37888" "64 18944"; do
  set -- $cfg
  timeout 600 python tools/affadd_k.py $1 $2 > gpurun_out/r2p_affadd_$1_$2.json 2> gpurun_out/r2p_affadd_$1_$2.err
  tail -4 gpurun_out/r2p_affadd_$1_$2.err
done
