# chunked fix-ups of the splice kernel: rounds every k groups of 2 KB (0 = after the whole gamete), configs[3] shapes and the headline
for f in 0 1 2 4 8; do echo "GSB200_SPLICE_FIX=$f"; GSB200_SPLICE_FIX=$f python profiles/scripts/config4_probe.py 500000 100000 both; done
