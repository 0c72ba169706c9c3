#pragma once
#include "../enoki_dyn.h"
