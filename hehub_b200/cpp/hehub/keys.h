// keys.h — RlweKsk, get_relin_key / get_conj_key / get_rot_key, RotKey (src/fhe/primitives/keys.{h,cpp}).
// The definitions live next to the key-switch container in rgsw.h; this header exists so that code written
// against the reference's include layout (`#include "fhe/primitives/keys.h"`) finds the same names.
#pragma once
#include "rgsw.h"
