#!/bin/bash
# One committed attempt to obtain GPy (the un-vendored dependency that carries the reference's GP arithmetic,
# /root/reference/requirements.txt:6) on a box: VERDICT r01 "Next round" item 1d.  Run through gpurun; the log is
# copied to profiles/r02_gpy_attempt.log.
out=${1:-gpurun_out/r02_gpy_attempt.log}
{
  echo "== date: $(date -u)"; echo "== host: $(hostname)"
  echo "== python -c 'import GPy, paramz'"; python -c 'import GPy, paramz; print(GPy.__version__, paramz.__version__)' 2>&1 | tail -2
  echo "== pip list | grep -i -E 'gpy|paramz'"; python -m pip list 2>/dev/null | grep -i -E 'gpy|paramz' || echo "(none)"
  echo "== wheelhouse"; ls /opt/wheelhouse 2>/dev/null | grep -i -E 'gpy|paramz' || echo "(no GPy / paramz wheel in /opt/wheelhouse)"
  echo "== find / -iname 'GPy*' -o -iname 'paramz*' (site-packages, caches)"
  find / -xdev \( -iname 'GPy' -o -iname 'GPy-*' -o -iname 'paramz*' \) -not -path '/proc/*' 2>/dev/null | head -5 || true
  echo "== pip install --no-index --find-links /opt/wheelhouse GPy==1.9.9 paramz"
  timeout 60 python -m pip install --no-index --find-links /opt/wheelhouse --target /tmp/gpy_try GPy==1.9.9 paramz 2>&1 | tail -4
  echo "== pip download GPy==1.9.9 (index; expected to fail: no network)"
  timeout 40 python -m pip download --no-deps -d /tmp/gpy_dl GPy==1.9.9 2>&1 | tail -4
  echo "== result: $(PYTHONPATH=/tmp/gpy_try python -c 'import GPy; print("GPy importable", GPy.__version__)' 2>&1 | tail -1)"
} > "$out" 2>&1
cat "$out"
