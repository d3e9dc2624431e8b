M2S_LIB=build/libm2s_ideal.so python scripts/_ideal.py
