"""C++ host mirror of the reference's KmerSpectrum / ReadSet / ReadSelector / FilterReads surface above the C ABI."""
