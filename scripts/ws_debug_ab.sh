for f in "-DCOUP_WS_DEBUG"; do
  COUP_B200_NVCC_EXTRA="$f" python -m open_spiel_coup_b200.build --force > /dev/null
  echo "== $f"; python scripts/ws_debug_probe.py
done
python -m open_spiel_coup_b200.build --force > /dev/null
