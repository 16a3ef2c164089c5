// oracle_more.hpp — TEST INFRASTRUCTURE, NOT PRODUCT CODE (see oracle_core.hpp header).
// Dispatch hook for the remaining registry models (Ising PT, PG logistic, variable selection).
#pragma once
#include "oracle_core.hpp"
#include "oracle_ising.hpp"
#include "oracle_pg.hpp"
#include "oracle_varsel.hpp"
#include "oracle_blasso.hpp"

#define ORC_MORE_DISPATCH(CALL)                                                                      \
  case UBM_MODEL_ISING_PT: { orc::IsingModel M((const ubm_model_ising_pt*)model); return CALL; }       \
  case UBM_MODEL_PG_LOGISTIC: { orc::PgLogisticModel M((const ubm_model_pg_logistic*)model); return CALL; } \
  case UBM_MODEL_VARSEL: { orc::VarselModel M((const ubm_model_varsel*)model); return CALL; }       \
  case UBM_MODEL_BLASSO: { orc::BlassoModel M((const ubm_model_blasso*)model); return CALL; }
