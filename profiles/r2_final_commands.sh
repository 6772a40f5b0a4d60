set -x
mkdir -p gpurun_out
python -m pytest tests -m gpu -q > gpurun_out/r2_final_pytest.log 2>&1; echo "pytest rc $?"
python -c 'import __graft_entry__ as g; g.smoke(); print("smoke ok")' > gpurun_out/r2_final_smoke.log 2>&1; echo "smoke rc $?"
python bench.py --steps 20 --warmup 5 > gpurun_out/r2_final_bench.json 2> gpurun_out/r2_final_bench.err; echo "bench rc $?"
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r2_final_bench_reference.json 2> gpurun_out/r2_final_bench_reference.err; echo "ref rc $?"
python profiles/trace_e2e.py > gpurun_out/r2_trace3.log 2>&1; echo "trace rc $?"
python profiles/trace_timeline.py > gpurun_out/r2_timeline_final.log 2>&1; python profiles/trace_chunks.py > gpurun_out/r2_chunks_final.log 2>&1; echo "timeline rc $?"
python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-extras > gpurun_out/r2_final_bench_short.json 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -c 6000 --csv --log-file gpurun_out/r2_final_launches_ncu.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-extras > gpurun_out/r2_final_ncu_bench.log 2>&1; echo "launchlist rc $?"
python profiles/profile_run.py 1.0 > gpurun_out/r2_final_prof_plain.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2_final_prof_launches.csv python profiles/profile_run.py 1.0 > gpurun_out/r2_final_prof_l.log 2>&1 && timeout 1500 ncu --set full --import-source on --clock-control none -k 'regex:k_pcr|k_fragment|k_intervals|k_prefix|k_expand|k_apply_picks|k_ingest|k_first_flags|k_emit|k_target|k_cut_index|k_copy_meta' -f -o gpurun_out/r2_final_full python profiles/profile_run.py 1.0 > gpurun_out/r2_final_prof_f.log 2>&1; echo "full rc $?"
ncu -i gpurun_out/r2_final_full.ncu-rep --page raw --csv > gpurun_out/r2_final_full_raw.csv 2>/dev/null; ls -la gpurun_out | tail -20
