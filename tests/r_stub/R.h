/* see Rinternals.h in this directory */
#ifndef RESNMTF_R_STUB_R_H
#define RESNMTF_R_STUB_R_H
#include <stdio.h>
#include <stdlib.h>
#endif
