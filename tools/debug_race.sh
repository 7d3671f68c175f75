for i in 1 2 3; do timeout 120 python tools/debug_disc.py 2>&1 | grep iter; done
for i in 1 2; do T2H_PDL=0 timeout 120 python tools/debug_disc.py 2>&1 | grep iter; done
for i in 1 2; do timeout 120 python tools/debug_disc.py sync 2>&1 | grep iter; done
