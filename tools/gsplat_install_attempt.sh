#!/bin/bash
# try to obtain the real gsplat==1.4.0 on the GPU box (the reference's arithmetic lives in that
# un-vendored wheel, /root/reference/requirements.txt:1).  The box has no network; every avenue is tried and logged.
# Output: gpurun_out/r02_gsplat_install_attempt.log (copied to profiles/).
LOG=gpurun_out/r02_gsplat_install_attempt.log
mkdir -p gpurun_out baseline/_ref
{
  echo "# gsplat==1.4.0 install attempt on the GPU box: $(date -u +%FT%TZ)"
  echo "## environment"; nvidia-smi --query-gpu=name,driver_version --format=csv,noheader | head -1; python --version; nvcc --version | tail -2
  echo "## 1. already importable?"; python -c "import gsplat; print('gsplat', gsplat.__version__, gsplat.__file__)" 2>&1 | tail -1
  echo "## 2. any copy on disk? (wheels, sdists, source trees)"
  find / -xdev \( -iname 'gsplat*' -o -iname 'nerfacc*' \) -not -path '/proc/*' -not -path "$PWD/*" 2>/dev/null | head -20
  ls /opt/wheelhouse 2>/dev/null | grep -i -E 'gsplat|nerfacc|jaxtyping' ; echo "(wheelhouse entries matching: $(ls /opt/wheelhouse 2>/dev/null | grep -c -i -E 'gsplat|nerfacc'))"
  echo "## 3. offline wheelhouse install"
  timeout 120 python -m pip install --no-index --no-build-isolation --find-links /opt/wheelhouse --target baseline/_ref gsplat==1.4.0 2>&1 | tail -4
  echo "## 4. index install (needs network)"
  timeout 60 python -m pip install --target baseline/_ref --no-deps gsplat==1.4.0 2>&1 | tail -4
  echo "## 5. pip download of the sdist (needs network)"
  timeout 60 python -m pip download --no-deps --no-binary :all: gsplat==1.4.0 -d /tmp/gsplat_sdist 2>&1 | tail -3
  echo "## 6. direct fetch from PyPI / GitHub (needs network)"
  timeout 20 curl -sS -I https://pypi.org/simple/gsplat/ 2>&1 | head -2
  timeout 20 curl -sS -I https://github.com/nerfstudio-project/gsplat/archive/refs/tags/v1.4.0.tar.gz 2>&1 | head -2
  timeout 20 git ls-remote https://github.com/nerfstudio-project/gsplat.git v1.4.0 2>&1 | tail -1
  echo "## result"; python -c "
import sys; sys.path.insert(0, 'baseline/_ref')
try:
    import gsplat; print('INSTALLED', gsplat.__version__)
except Exception as e:
    print('NOT AVAILABLE:', type(e).__name__, e)"
} > $LOG 2>&1
tail -3 $LOG
