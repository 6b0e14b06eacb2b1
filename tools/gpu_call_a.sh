#!/bin/bash
# round-2 GPU call A: baseline sanity, flake hunt, MMA floors, TF32 peak, real callers (ref vs ours)
set -u
cd "$(dirname "$0")/.."
O=gpurun_out
mkdir -p $O
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $O/a_smi.txt 2>&1
( timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 ) > $O/a_pytest.log
( timeout 120 ./tools/mma_floor ) > $O/a_mma_floor.txt 2>&1
( timeout 120 python tools/tf32_peak.py ) > $O/a_tf32_peak.json 2> $O/a_tf32_peak.err
( timeout 600 python tools/flake_hunt.py --iters 300 ) > $O/a_flake_tc.txt 2>&1
( PN2_TC=0 timeout 600 python tools/flake_hunt.py --iters 300 ) > $O/a_flake_ffma.txt 2>&1
( timeout 600 python tools/flake_hunt.py --iters 300 --interleave 0 ) > $O/a_flake_tc_nointer.txt 2>&1
T=/tmp/rc; mkdir -p $T
( timeout 600 python tests/ref_callers.py --impl ref --case backbone_train --out $O/a_bb_ref.npz --save-state $T/bb.pt ) > $O/a_bb_ref.log 2>&1
( timeout 600 python tests/ref_callers.py --impl ours --case backbone_train --out $O/a_bb_ours.npz --load-state $T/bb.pt ) > $O/a_bb_ours.log 2>&1
( timeout 900 python tests/ref_callers.py --impl ref --case pq_eval --batch 8 --out $O/a_pq_ref.npz --save-state $T/pq.pt ) > $O/a_pq_ref.log 2>&1
( timeout 900 python tests/ref_callers.py --impl ours --case pq_eval --batch 8 --out $O/a_pq_ours.npz --load-state $T/pq.pt ) > $O/a_pq_ours.log 2>&1
( timeout 600 compute-sanitizer --tool racecheck --print-limit 20 python -m pytest tests/test_gpu_fused.py -m gpu -q -x -k "fp_matches" 2>&1 | tail -60 ) > $O/a_racecheck.txt
( timeout 600 compute-sanitizer --tool initcheck --print-limit 20 python -m pytest tests/test_gpu_fused.py -m gpu -q -x -k "fp_matches" 2>&1 | tail -60 ) > $O/a_initcheck.txt
echo done
