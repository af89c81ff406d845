/* oracle/ref_model: TEST INFRASTRUCTURE.  Stand-in for the CGNS library header of this name: the NS interface sources include the CGNS viewer header without using CGNS */
#pragma once
#define CGNS_ENUMT(t) int
const char *cg_get_error(void);
