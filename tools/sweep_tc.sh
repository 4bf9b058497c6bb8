for cap in 96 128; do
echo "== BNCAP=$cap"; PCODEC_TC_BNCAP=$cap MODES=0 timeout 200 python tools/exp_tc.py 2>&1 | cut -c1-80
done
