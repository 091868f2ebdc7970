python tools/tc_probe.py --cases small 2>&1 | tail -5
python tools/tc_probe.py --cases sweep 2>&1 | tail -30
for ns in 8 16 32; do
 for dm in 0 1 2; do
  echo "== ns=$ns debug_mode=$dm"
  python tools/kbench.py --packed --streams 64 --samples '2**24' --option variant=13 --option tc_ns=$ns --option debug_mode=$dm --dbg 2>&1 | tail -5
 done
done
