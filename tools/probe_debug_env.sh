T='import torch; torch.cuda.set_device(0); print(torch.zeros(1,device="cuda").item())'
echo "--- A coredump only"; CUDA_ENABLE_COREDUMP_ON_EXCEPTION=1 CUDA_COREDUMP_FILE=/tmp/core_a python -c "$T" 2>&1 | tail -2
echo "--- B lightweight"; CUDA_ENABLE_COREDUMP_ON_EXCEPTION=1 CUDA_ENABLE_LIGHTWEIGHT_COREDUMP=1 CUDA_COREDUMP_FILE=/tmp/core_b python -c "$T" 2>&1 | tail -2
echo "--- C user-triggered"; CUDA_ENABLE_USER_TRIGGERED_COREDUMP=1 python -c "$T" 2>&1 | tail -2
echo "--- D cuda-gdb"; timeout 120 cuda-gdb -batch -ex run -ex bt --args python -c "$T" 2>&1 | tail -5
echo "--- E sanitizer"; timeout 300 compute-sanitizer --tool memcheck python -c "$T" 2>&1 | tail -3
echo "--- dmesg"; dmesg 2>&1 | tail -3
nvidia-smi -q | grep -i -E "mig|compute mode|persistence" | head
