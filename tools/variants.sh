set -e
cd python_bulletproofs_b200
cp libbpgpu.so /tmp/lib_b4.so
python ../tools/msm_probe.py --lgn 18,20 --c 0 2>&1 | tail -2
for b in 5 6; do cp libbpgpu_b$b.so libbpgpu.so; echo "== minBlocks $b"; python ../tools/msm_probe.py --lgn 18,20 --c 0 2>&1 | tail -2; done
cp /tmp/lib_b4.so libbpgpu.so
