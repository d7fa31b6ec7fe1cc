for i in 1 2; do
echo "--- kat only run $i"; python -m pytest tests/test_fft_kat.py -m gpu -q -x -p no:faulthandler 2>&1 | tail -2
echo "--- lengths only run $i"; python -m pytest tests/test_gpu_parity.py -m gpu -q -x -p no:faulthandler -k "lengths or fft_golden or noncontig" 2>&1 | tail -2
done
echo "--- MALLOC_CHECK"; MALLOC_CHECK_=3 cuda-gdb -batch -ex "set pagination off" -ex run -ex bt --args python -m pytest tests -m gpu -q -x -p no:faulthandler -k "fft" 2>&1 | grep -v "^\[" | head -40
