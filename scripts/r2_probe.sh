#!/bin/bash
# round 2, first GPU call: (1) is any MuJoCo importable on the GPU box? (2) register-cap cost of k_env_step at fixed occupancy
cd "$GRAFT_REPO_ROOT"
{
echo "== mujoco probe"; 
python -c "import mujoco; print('mujoco', mujoco.__version__)" 2>&1 | tail -1
python -c "import sys; sys.path=[p for p in sys.path if p not in ('', '.', '$GRAFT_REPO_ROOT', '/root/repo')]; import mujoco_py; print('mujoco_py', mujoco_py.__file__)" 2>&1 | tail -1
python -c "import dm_control" 2>&1 | tail -1
timeout 60 python -m pip download mujoco -d /tmp/mj --no-deps 2>&1 | tail -2
ls /opt/wheelhouse 2>/dev/null | grep -i -E "mujoco|gym|dm_control|brax" ; echo "wheelhouse grep rc=$?"
ls /root/.mujoco 2>&1 | tail -1
nvidia-smi --query-gpu=name,clocks.max.sm,memory.total --format=csv,noheader
nproc; cat /sys/fs/cgroup/cpu.max
} > gpurun_out/r2_mujoco_probe.txt 2>&1
for v in 2 3 4; do
  echo "== variant min_ctas=$v" 
  UHC_B200_SO=$PWD/build_variants/lib_cta$v.so timeout 300 python scripts/quick_time.py 4096 20
done > gpurun_out/r2_regcap.txt 2>&1
cat gpurun_out/r2_mujoco_probe.txt gpurun_out/r2_regcap.txt
