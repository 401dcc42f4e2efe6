for span in 98304 131072 196608; do for chunk in 32768 49152 65536 131072; do
  if [ $chunk -le $span ]; then echo -n "span $span chunk $chunk: "; SPART_HOST_SPAN=$span SPART_HOST_CHUNK=$chunk python tools/e2e_trace.py 2>&1 | grep "ms per call"; fi
done; done
