#pragma once
#include "../enoki_scalar_stub.h"
