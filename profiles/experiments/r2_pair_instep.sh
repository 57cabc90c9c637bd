# in-step effect of running pass D2 as CTA pairs at C4 (alternating runs on one box)
for i in 1 2; do
for X in 0 1; do
RZ_EXP_PAIR=$X python bench.py --workload contrastive --steps 20 --warmup 5 --no-cpu --no-addons 2>/dev/null | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('pair=$X', round(d['ms_per_step'],2), d['loss'], d['grad_checksum'], d['clocks'])"
done; done
