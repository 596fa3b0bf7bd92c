# cache-policy builds of the direct-load kernel (build/alt/lib_ld<L>st<S>.so: -DNI_LOAD_POLICY=L -DNI_STORE_POLICY=S)
set -u
L=""
for n in ld1st0 ld0st0 ld1st1 ld0st1 ld3st0 ld3st1 ld4st0 ld4st1 ld6st1 ld5st1; do L="$L --lib build/alt/lib_$n.so"; done
timeout 400 python scripts/sweep.py --cfgs c2,c3 $L
timeout 100 python scripts/sweep.py --cfgs c3,c5 --lib build/alt/lib_ld1st0.so --lib build/alt/lib_ld0st1.so --lib build/alt/lib_ld0st0.so
