# warps-per-CTA sweep (tuning aid): bash tools/nwsweep.sh <config> <nw> [<nw> ...]
C=$1; shift
for NW in "$@"; do RFSB200_WARPS_PER_CTA=$NW python bench.py --config $C --steps 40 --warmup 5 --no-cpu-baseline --no-e2e 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('$C nw $NW kernel_us %.1f ms %.4f' % (d['roofline']['kernel_us'], d['ms_per_step']))"; done
