"""Drop-in for the reference's main.py (main.py:49-65): the FMIndex demo on "this is an example text" followed by
the wavelet-tree compression print-out, running on the B200 kernels.  The reference duplicates the FMIndex class
here (main.py:6-46 == csa/csa.py:6-45); this file imports it instead."""
from csa.csa import FMIndex
from csa.wavelet_tree import WaveletTree


def main():
    text = "this is an example text"
    fm_index = FMIndex(text)
    pattern = "example"
    print(f"Searching for the pattern: '{pattern}'")
    matches = fm_index.find_pattern(pattern)
    print(f"Pattern '{pattern}' found at indices: {matches}")
    wavelet_tree = WaveletTree(text)
    compressed = wavelet_tree.compress()
    print(f"Wavelet Tree Compression: {compressed}")


if __name__ == "__main__":
    main()
