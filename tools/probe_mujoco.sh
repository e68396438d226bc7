#!/bin/bash
# Round-2 probe: can the GPU box run the reference's own physics (mujoco==3.3.3)?  Output is committed as
# profiles/r2_mujoco_probe.log.
mkdir -p gpurun_out
{
  echo "== python imports"
  for mod in mujoco gymnasium stable_baselines3 imitation mujoco_mjx dm_control brax; do
    python -c "import $mod; print('$mod', getattr($mod, '__version__', '?'))" 2>&1 | tail -1
  done
  echo "== pip (no index reachable?)"
  timeout 60 python -m pip download mujoco==3.3.3 -d /tmp/w 2>&1 | tail -2
  timeout 60 python -m pip install --no-index --find-links /opt/wheelhouse --target /tmp/mj mujoco==3.3.3 2>&1 | tail -2
  echo "== wheelhouse / filesystem"
  ls /opt/wheelhouse 2>/dev/null | grep -i -E "mujoco|gymnasium|stable|imitation" || echo "no mujoco/gymnasium/sb3 wheel in /opt/wheelhouse"
  find / \( -iname "*mujoco*" -o -iname "libmujoco*" \) -not -path "/proc/*" -not -path "*/gpurun*" 2>/dev/null | grep -v -E "ur3e|SURVEY|probe" | head
  ls baseline/_ref 2>&1 | head -3
  echo "== box"
  nvidia-smi -L; nproc; free -g | head -2
} > gpurun_out/r2_mujoco_probe.log 2>&1
cat gpurun_out/r2_mujoco_probe.log
