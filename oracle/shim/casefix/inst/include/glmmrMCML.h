// oracle/shim/casefix — the reference's src/*.cpp include "../inst/include/glmmrMCML.h" while the file is inst/include/glmmrmcml.h
// (works on a case-insensitive file system only).  With -Ioracle/shim/casefix/src that include resolves to this file, which
// forwards to the reference's real header (found through -I<reference>/inst/include).  TEST INFRASTRUCTURE.
#pragma once
#include "glmmrmcml.h"
