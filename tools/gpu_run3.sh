for c in 4 3 2; do for s in 2 4 6; do STRGPU_PRE_CTAS=$c python bench.py --steps 3 --warmup 3 --no-cpu-baseline --streams $s 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('pre_ctas',$c,'streams',$s,'value',d['value'],'ms',d['ms_per_step'],'frac',d['roofline']['frac'])"; done; done
STRGPU_PRE_CTAS=3 STRGPU_MAX_STAGE=0 python tools/profile_scan.py | grep us_per
