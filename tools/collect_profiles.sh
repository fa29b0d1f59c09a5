#!/bin/bash
# ncu evidence of round 2 (run from the repo root on a GPU box; every profiled command first runs once without ncu).
set -x
O=gpurun_out
python bench.py --steps 2 --warmup 3 --no-cpu > $O/r02_bench_plain.json 2> $O/r02_bench_plain.err || exit 1
ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file $O/r02_bench_launches_raw.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu > $O/r02_bench_ncu.log 2>&1
python tools/prof_driver.py qp 4096 > $O/r02_qp_plain.log 2>&1 || exit 1
ncu --set full --clock-control none --import-source on -k regex:qp_kkt_sqd -s 1 -c 1 -o $O/r02_qp_sqd -f python tools/prof_driver.py qp 4096 > $O/r02_qp_ncu.log 2>&1
python tools/prof_driver.py sparse > $O/r02_sparse_plain.log 2>&1 || exit 1
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__warps_active.avg.pct_of_peak_sustained_active,smsp__issue_active.avg.pct_of_peak_sustained_active \
    --clock-control none -c 400 --csv --log-file $O/r02_sparse_launches_raw.csv python tools/prof_driver.py sparse > $O/r02_sparse_ncu.log 2>&1
python tools/prof_driver.py conic_batch 148 20 > $O/r02_conic_batch_plain.log 2>&1 || exit 1
ncu --set full --clock-control none --import-source on -k regex:lsqr_batch -c 1 -o $O/r02_lsqr_batch -f python tools/prof_driver.py conic_batch 148 20 > $O/r02_conic_batch_ncu.log 2>&1
ls -la $O | grep r02_
# later in round 2: the shape-generic LDL' kernel (n=100, m=50, p=0, 10 active) and the PSD route of config 5
python tools/prof_driver.py qp 4096 100 50 0 10 > $O/r02_qp_any_plain.log 2>&1 || exit 1
ncu --set full --clock-control none --import-source on -k regex:qp_kkt_sqd_any -s 1 -c 1 -o $O/r02_qp_sqd_any -f python tools/prof_driver.py qp 4096 100 50 0 10 > $O/r02_qp_any_ncu.log 2>&1
python bench_aux.py --configs 5 > $O/r02_psd_plain.log 2>&1 || exit 1
ncu --metrics gpu__time_duration.sum --clock-control none -c 40 --csv --log-file $O/r02_psd_launches_raw.csv python bench_aux.py --configs 5 > $O/r02_psd_ncu.log 2>&1
python bench_aux.py --configs 1,2s,3,3p,4,4c,4x,5 > $O/r02_bench_aux.jsonl 2> $O/r02_bench_aux.err
ls -la $O | grep r02_
# end of round 2: sparse path after the row-ordered launch lists / asynchronous tiles / assembly nodes
python tools/prof_driver.py sparse > $O/r02_sparse_plain.log 2>&1 || exit 1
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__warps_active.avg.pct_of_peak_sustained_active,smsp__issue_active.avg.pct_of_peak_sustained_active \
    --clock-control none -c 400 --csv --log-file $O/r02_sparse_launches_raw.csv python tools/prof_driver.py sparse > $O/r02_sparse_ncu.log 2>&1
python tools/sparse_launch_report.py $O/r02_sparse_launches_raw.csv > $O/r02_sparse_launches.txt
ncu --set full --clock-control none --import-source on -k regex:"mf_(forward|backward|factor)" -s 34 -c 34 -o $O/r02_sparse_kernels -f python tools/prof_driver.py sparse > $O/r02_sparse_kernels_ncu.log 2>&1
python bench_aux.py --configs 3pnocpu > $O/r02_sparse_portfolio_plain.log 2>&1 || exit 1
ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file $O/r02_sparse_portfolio_launches_raw.csv python bench_aux.py --configs 3pnocpu > /dev/null 2>&1
