cd "${GRAFT_REPO_ROOT:-.}"
for b in 16 32 64; do
timeout 300 python tools/fits_scaling.py 256 512 --batch=$b 2>&1 | tail -2 | cut -c1-330
done
