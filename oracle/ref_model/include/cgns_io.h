/* oracle/ref_model: TEST INFRASTRUCTURE.  Stand-in for the CGNS library header of this name (nothing of it is used) */
#pragma once
