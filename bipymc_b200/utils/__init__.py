"""Import-path mirror of bipymc/utils (banana_rv, dblgauss_rv, d100_gauss)."""
