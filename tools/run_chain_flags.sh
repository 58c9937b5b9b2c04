# bring-up: chain kernel phase timelines with parts switched off (MVAE_CHAIN_DEBUG bits; results are garbage)
for f in ${FLAGS:-0 64 128}; do
  echo "=== flags $f"
  MVAE_CHAIN_DEBUG=$f timeout 120 python tools/chain_timeline.py 2>&1 | grep -v "begin"
done
