#pragma weak lamsa_aln_core
