python bench.py --no-cpu-baseline --headline-only 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); b=d['breakdown']; print('train', d['ms_per_step'], b['cednerf_march'])"
python profiles/tools/exp_render.py 2>&1 | grep "ms/frame\|cednerf_march_round "
