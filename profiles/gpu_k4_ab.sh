# K4 A/B: unstaged / staged with 256-, 128-, 64-frame chunks, on the C2 shape (level 5) and on a C3-like shape (8000 utterances, level 13)
cd $GRAFT_REPO_ROOT
cp webspeechanalyzer_b200/libfa_b200.so /tmp/orig.so
for v in nostage k256 k128 k64; do
  cp webspeechanalyzer_b200/build/variants/lib_$v.so webspeechanalyzer_b200/libfa_b200.so
  a=$(timeout 200 python profiles/stage_times.py $v= 2>/dev/null | tail -1 | python -c "import json,sys; print(json.loads(sys.stdin.read())['features'])")
  b=$(N_UTT=8000 LEVEL=13 WANT_SPEC=0 timeout 300 python profiles/stage_times.py $v= 2>/dev/null | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(d['features'], d['feature_rows'])")
  echo "$v: C2 level 5 features $a ms; 8000 utt level 13 features $b"
done
cp /tmp/orig.so webspeechanalyzer_b200/libfa_b200.so
