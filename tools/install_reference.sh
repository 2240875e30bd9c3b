#!/bin/bash
# Puts the UNMODIFIED reference into baseline/_ref (git-ignored; travels to the GPU box with gpurun) for
# tools/ref_gpu_bench.py.  The contract's pip recipe installs only the directories that carry an __init__.py
# (setup.py uses find_packages(); inferix/pipeline/self_forcing, .../causvid etc. are namespace packages and are
# skipped), so the files it left out are added from the same source tree without overwriting anything.
set -e
cd "$(dirname "$0")/.."
rm -rf /tmp/refcopy && cp -r /root/reference /tmp/refcopy      # the build writes into the source tree; /root/reference is read-only
python -m pip install --no-index --no-build-isolation --no-deps --find-links /opt/wheelhouse --target baseline/_ref --upgrade /tmp/refcopy
cp -rn /root/reference/inferix/. baseline/_ref/inferix/
find baseline/_ref -name __pycache__ -type d -prune -exec rm -rf {} +
echo "installed: $(find baseline/_ref/inferix -name '*.py' | wc -l) python files"
