#!/bin/bash
# Runs UNCHANGED drivers of the reference against this engine (drop-in check, GPU box).
#   REF=/path/to/ADI_thermal_fields tools/run_reference_drivers.sh [tag]
# Drivers are started with `python -m <module>` from a scratch directory: started as `python $REF/x.py`
# the script's own directory comes first on sys.path and the reference's adi3d_gpu_coeff.py would
# shadow the drop-in (it then runs its CuPy algorithm on the `cupy` shim -- correct, but not this engine).
# The reference tree is not part of this repo and is not on the GPU box by default; for a one-off
# run put a scratch copy under the git-ignored baseline/_ref/ and delete it afterwards.
tag=${1:-drv}
REF=${REF:-baseline/_ref}
ROOT=$(cd "$(dirname "$0")/.." && pwd)
REF=$(cd "$ROOT" && cd "$REF" && pwd)
OUT=$ROOT/gpurun_out; mkdir -p $OUT
export NUMBA_CACHE_DIR=$(mktemp -d) PYTHONDONTWRITEBYTECODE=1
export PYTHONPATH=$ROOT/adi_thermal_fields_b200/dropin:$ROOT/tools/stubs:$REF
cd $(mktemp -d)
{
echo "== quick_compare_neumann_robin_backend.py --backend both (Cartesian, CPU numba vs GPU engine)"
python -m quick_compare_neumann_robin_backend --backend both --nxr 24 --nz 256 --nframes 4 --tmax 60 2>&1 | grep -E "^\[|Error|error|Traceback" 
echo "== quick_compare_layer_birth_robin_v3.py --backend cpu / gpu (Cartesian layer births)"
for b in cpu gpu; do python -m quick_compare_layer_birth_robin_v3 --backend $b --nxr 16 --N_total 3 --nframes 5 --d 0.02 --z_back 0.1 --z_front 0.08 --t_tail 40 2>&1 | grep -E "^\[|Error|error|Traceback" | tail -12; done
echo "== quick_compare_layer_birth_robin_cyl_v3.py (cylindrical nz-growing births, GPU engine behind adi3d_cyl_phi_v3)"
python -m quick_compare_layer_birth_robin_cyl_v3 --R 0.02 --z_back 0.02 --d 0.005 --t_step 0.5 --N_total 3 --t_tail 0.5 --nr 16 --nphi 32 --h_side 500 --h_end 500 --T_inf 20 --Ts 1000 --nframes 5 2>&1 | grep -vi "warn" | tail -12
echo "== tests/test_spiral_vs_analytic.py (the reference's only test; numerics on the GPU engine)"
python -c "import adi3d_gpu_coeff, adi3d_cyl_phi_v3, quick_spiral_deposition_gif_v5 as q; print('[modules]', adi3d_gpu_coeff.__file__, adi3d_cyl_phi_v3.__file__, q.__file__)"
python -m pytest $REF/tests/test_spiral_vs_analytic.py -x -q -p no:cacheprovider --rootdir=$PWD 2>&1 | tail -15
} | tee $OUT/${tag}_drivers.txt
