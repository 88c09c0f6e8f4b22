#pragma weak lamsa_res_aux
