set -x
cd $GRAFT_REPO_ROOT
timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -6 > gpurun_out/final_pytest.log
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/final_smoke.log 2>&1
python bench.py --steps 20 --warmup 5 > gpurun_out/final_bench_n1.json 2> gpurun_out/final_bench_n1.err
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/final_bench_ref.json 2> gpurun_out/final_bench_ref.err
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/final_train_launches.csv python tools/prof_train.py 256 256 2 > gpurun_out/final_ncu1.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/final_train_b32_launches.csv python tools/prof_train.py 32 256 2 > gpurun_out/final_ncu2.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:ctc_loss_fast -s 3 -c 1 -o gpurun_out/final_ctc_bw python tools/prof_ctc.py 65536 > gpurun_out/final_ncu3.log 2>&1
ncu --set full --clock-control none -k regex:ctc_loss_fast -s 3 -c 1 -o gpurun_out/final_ctc_cfg2 python tools/prof_ctc.py 256 > gpurun_out/final_ncu4.log 2>&1
ls -la gpurun_out | tail -20
