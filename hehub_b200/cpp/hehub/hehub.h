// hehub.h — umbrella header of the B200-native host mirror of the hehub:: ciphertext-op API.
#pragma once
#include "bgv.h"
#include "ckks.h"
#include "mod_arith.h"
#include "ntt.h"
#include "permutation.h"
#include "rgsw.h"
#include "rlwe.h"
#include "rns.h"
#include "rns_transform.h"
#include "serialize.h"
