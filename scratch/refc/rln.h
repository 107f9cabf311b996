#include "rln_b200.h"
