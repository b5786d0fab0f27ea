/* empty shim: amd.h is only referenced on the non-METIS branch (LSparsity.h:615-620) */
