set -e
cd "$(dirname "$0")/.."
OUT=profiles/r02n_sass_scan_loop.txt
{
echo "# SASS evidence for the final round-2 build (cuobjdump -sass of the sm_100a objects, produced in the build container by"
echo "# scripts/sass_evidence.sh; no GPU needed).  (1) tensor-memory / tcgen05 mnemonics per object, (2) the instruction mix and the"
echo "# full listing of the LAP scan loop of ged_solve_kernel<2, tmem> (the n <= 64 class, 53 % of the benchmarked step's instructions)."
echo
echo "## tcgen05 / tensor-memory mnemonics (count per object)"
for o in ppo-bihyb_b200/build/ged_s*.o ppo-bihyb_b200/build/ged_team.o ppo-bihyb_b200/build/ged_main.o; do
  printf "%s: " "$(basename $o)"
  cuobjdump -sass $o | grep -oE "\b(LDTM|STTM|UTCBAR|UTCATOMSWS|UTCMMA|UTCHMMA|UTMALDG|CREDUX|REDUX)[A-Za-z0-9_.]*" | sed 's/\..*//' | sort | uniq -c | awk '{printf "%s=%s ", $2, $1}'
  echo
done
echo
echo "## scan loop, ged_solve_kernel<2, kModeTmem>"
python scripts/sass_scan_loop.py ppo-bihyb_b200/build/ged_s2.o "ged_solve_kernelILi2ELi1E" --dump | awk "/^loop 0x1/{exit} {print}"
echo
echo "## scan loop instruction mix, other instantiations (first scan loop of each; mode 0 shared, 1 tensor, 2 tensor + shared, 3 tensor + global memory)"
for s in 1 3 4; do for m in 0 1 2 3; do
  python scripts/sass_scan_loop.py ppo-bihyb_b200/build/ged_s$s.o "ged_solve_kernelILi${s}ELi${m}E" 2>/dev/null | sed -n 2,3p | sed "s/^/S=$s mode=$m  /"
done; done
echo "team W=4:"; python scripts/sass_scan_loop.py ppo-bihyb_b200/build/ged_team.o "ged_solve_team_kernelILi4ELi4E" 2>/dev/null | sed -n 2,3p
} > $OUT
wc -l $OUT; head -20 $OUT
