cd "${GRAFT_REPO_ROOT:-.}"
for v in 0 1; do
if [ $v = 1 ]; then export CS_SINGLE_PANEL=1; else unset CS_SINGLE_PANEL; fi
echo "single_panel_for_batches=$v"
timeout 500 python tools/batch_phases.py 2>&1 | grep "tile= 64" | grep "B=32\|B=64"
timeout 300 python tools/fits_scaling.py 256 512 --batch=32 2>&1 | tail -2 | cut -c1-200
done
