cd /tmp && rm -rf sd && mkdir sd && cd sd
export EMCGPU_SHARD=1 WORLD_SIZE=2 EMCNCCL_ID_FILE=/tmp/sd/id
for rb in 1 0; do
for r in 0 1; do mkdir -p rb$rb/r$r; (cd rb$rb/r$r; RANK=$r LOCAL_RANK=$r EMCNCCL_ID_FILE=/tmp/sd/id$rb /root/repo/viennaemc_b200/bin/resistor2D --seed 5 --steps 400 --transient 100 --avg 100 --red-black $rb > out.txt 2>&1) & done; wait
echo "red-black $rb:"; tail -n 4 rb$rb/r0/out.txt; tail -n 2 rb$rb/r1/out.txt
done
