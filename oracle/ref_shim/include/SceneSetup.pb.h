/* TEST INFRASTRUCTURE (oracle/_ref build only): see Vector.pb.h */
#pragma once
#include "Vector.pb.h"
