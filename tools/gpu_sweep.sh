cd "${GRAFT_REPO_ROOT:-.}"
for w in 4 6 8 11 12; do
echo "== CS_PANEL=$w"
CS_PANEL=$w timeout 300 python tools/batch_phases.py 2>&1 | grep "tile= 64" | grep "B= 1 \|B=64"
CS_PANEL=$w timeout 300 python tools/fits_scaling.py 256 --batch=32 2>&1 | tail -1 | cut -c1-200
done
