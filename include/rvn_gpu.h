/* rvn_gpu.h -- thin C ABI between Raven-surface host code and the sm_100a CUDA engine.
 *
 * This is the drop-in boundary of the B200 build.  The reference has no
 * binary plugin loader: its hot path is the free function
 *     void MassEnergyBalance(CModel*, const optStruct&, const time_struct&)   (reference src/Solvers.cpp:19-21)
 * fed by CModel::UpdateHRUForcingFunctions (src/UpdateForcings.cpp:30) and followed by
 * IncrementCumulInput/IncrementCumOutflow (src/Model.cpp:2458,2493) and the WriteMinorOutput reductions
 * (src/StandardOutput.cpp:745-785).  The host side (ravenhydroframework_b200/host, C++) keeps those
 * signatures and flattens the model it parsed from .rvi/.rvh/.rvp/.rvt/.rvc into the plain-C descriptor
 * below; everything that happens per timestep then runs on the GPU behind these entry points.
 *
 * Conventions: plain pointers + sizes, no C++/torch types; every function returns 0 on success or a
 * negative rvn_status and leaves a message readable through rvn_gpu_last_error().  A handle owns one
 * CUDA stream and is thread-compatible (not thread-safe), like the reference (one model per process).
 * All floating point on this path is IEEE FP64.  There is NO CPU fallback: if no CUDA device is present
 * rvn_gpu_create fails with RVN_ERR_NO_DEVICE.
 */
#ifndef RVN_GPU_H
#define RVN_GPU_H
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

#define RVN_ABI_VERSION 11
#define RVN_DOESNT_EXIST (-1)
#define RVN_CONV_STORES 50   /* MAX_CONVOL_STORES, reference src/RavenInclude.h / Convolution.cpp:17 */
#define RVN_MAX_CONN 24      /* max connections of a non-convolution process (Precipitation: 13..19) */
#define RVN_MAX_SOIL_LAYERS 8
#define RVN_MAX_RIVER_SEGS 50 /* MAX_RIVER_SEGS, reference src/SubBasin.h */
#define RVN_RES_CURVES 5      /* stage [m], weir/curve outflow Q, underflow Q [m3/s], area [m2], volume [m3] */
#define RVN_RES_STATE 8       /* doubles of state per (member, reservoir) */

typedef enum {
  RVN_OK = 0, RVN_ERR_NO_DEVICE = -1, RVN_ERR_BAD_ARG = -2, RVN_ERR_CUDA = -3, RVN_ERR_UNSUPPORTED = -4,
  RVN_ERR_NOMEM = -5
} rvn_status;

/* ---- state-variable types: numbering == reference enum sv_type (src/RavenInclude.h:876-951) so that the
 *      SV catalogue can be compared as integers against the reference build. ---- */
typedef enum {
  RVN_SV_SURFACE_WATER = 0, RVN_SV_ATMOSPHERE = 1, RVN_SV_ATMOS_PRECIP = 2, RVN_SV_PONDED_WATER = 3,
  RVN_SV_SOIL = 4, RVN_SV_CANOPY = 5, RVN_SV_CANOPY_SNOW = 6, RVN_SV_TRUNK = 7, RVN_SV_ROOT = 8,
  RVN_SV_GROUNDWATER = 9, RVN_SV_DEPRESSION = 10, RVN_SV_SNOW = 11, RVN_SV_NEW_SNOW = 12, RVN_SV_SNOW_LIQ = 13,
  RVN_SV_TOTAL_SWE = 14, RVN_SV_WETLAND = 15, RVN_SV_GLACIER = 16, RVN_SV_GLACIER_ICE = 17, RVN_SV_FIRN = 18,
  RVN_SV_LAKE_STORAGE = 19, RVN_SV_GLACIER_MB = 20, RVN_SV_CONVOLUTION = 21, RVN_SV_CONV_STOR = 22,
  RVN_SV_MIN_DEP_DEFICIT = 23, RVN_SV_CONTRIB_FRAC = 24, RVN_SV_INFIL_CORR = 25, RVN_SV_CUM_INFIL = 26,
  RVN_SV_GA_MOISTURE_INIT = 27, RVN_SV_CUM_SNOWMELT = 28, RVN_SV_AET = 29, RVN_SV_RUNOFF = 30,
  RVN_SV_ENERGY_LOSSES = 31, RVN_SV_SURFACE_WATER_TEMP = 32, RVN_SV_SNOW_TEMP = 33, RVN_SV_COLD_CONTENT = 34,
  RVN_SV_SOIL_TEMP = 35, RVN_SV_CANOPY_TEMP = 36, RVN_SV_SOIL_PCT_FROZ = 37, RVN_SV_SNOW_DEPTH = 38,
  RVN_SV_FIRN_GRAVITY = 39, RVN_SV_PERMAFROST_DEPTH = 40, RVN_SV_THAW_DEPTH = 41, RVN_SV_SNOW_DEPTH_STDDEV = 42,
  RVN_SV_SNOW_COVER = 43, RVN_SV_GLACIER_CC = 44, RVN_SV_SNOW_DEFICIT = 45, RVN_SV_SNOW_AGE = 46,
  RVN_SV_SNODRIFT_TEMP = 47, RVN_SV_SNOW_DRIFT = 48, RVN_SV_ICE_THICKNESS = 49, RVN_SV_SNOW_ALBEDO = 50, RVN_SV_CROP_HEAT_UNITS = 51,
  RVN_SV_NTYPES = 64
} rvn_sv_type;

/* HRU types: numbering == reference enum HRU_type (src/RavenInclude.h:599-609) */
typedef enum {
  RVN_HRU_STANDARD = 0, RVN_HRU_LAKE = 1, RVN_HRU_WATER = 2, RVN_HRU_GLACIER = 3, RVN_HRU_WETLAND = 4,
  RVN_HRU_MASKED_GLACIER = 5, RVN_HRU_ROCK = 6
} rvn_hru_type;

/* ---- process opcodes: one per (process class, algorithm) the device interpreter implements.
 *      Each cites the reference GetRatesOfChange/ApplyConstraints it replaces. ---- */
typedef enum {
  RVN_OP_NOP = 0,
  RVN_OP_PRECIPITATION,        /* CmvPrecipitation        src/PartitionPrecip.cpp:135-358          */
  RVN_OP_SNOBAL_HMETS,         /* CmvSnowBalance          src/SnowBalance.cpp:472-522              */
  RVN_OP_SNOBAL_SIMPLE_MELT,   /* CmvSnowBalance          src/SnowBalance.cpp:393-396,1311-1317    */
  RVN_OP_SNOBAL_CEMA_NEIGE,    /* CmvSnowBalance          src/SnowBalance.cpp:398-413              */
  RVN_OP_SNOBAL_HBV,           /* CmvSnowBalance          src/SnowBalance.cpp:430-465              */
  RVN_OP_SNOWREFREEZE_DD,      /* CmvSnowRefreeze         src/SnowMeltRefreeze.cpp:324-365         */
  RVN_OP_SNOTEMP_NEWTONS,      /* CmvSnowTempEvolve       src/SnowTempEvolve.cpp:42-77             */
  RVN_OP_INF_HMETS,            /* CmvInfiltration         src/Infiltration.cpp:491-514,605-626     */
  RVN_OP_INF_HBV,              /* CmvInfiltration         src/Infiltration.cpp:384-398             */
  RVN_OP_INF_GR4J,             /* CmvInfiltration         src/Infiltration.cpp:477-489             */
  RVN_OP_OVERFLOW,             /* CmvOverflow             src/Flush.cpp:259-295                    */
  RVN_OP_FLUSH,                /* CmvFlush                src/Flush.cpp:75-110                     */
  RVN_OP_SPLIT,                /* CmvSplit                src/Flush.cpp:178-210                    */
  RVN_OP_BASE_LINEAR,          /* CmvBaseflow             src/Baseflow.cpp:180-186,297-310         */
  RVN_OP_BASE_POWER_LAW,       /* CmvBaseflow             src/Baseflow.cpp                          */
  RVN_OP_BASE_GR4J,            /* CmvBaseflow             src/Baseflow.cpp                          */
  RVN_OP_PERC_LINEAR,          /* CmvPercolation          src/Percolation.cpp:237-243,339-360      */
  RVN_OP_PERC_CONSTANT,        /* CmvPercolation          src/Percolation.cpp:199-202              */
  RVN_OP_PERC_GR4J,            /* CmvPercolation          src/Percolation.cpp                       */
  RVN_OP_PERC_GR4JEXCH,        /* CmvPercolation          src/Percolation.cpp                       */
  RVN_OP_PERC_GR4JEXCH2,       /* CmvPercolation          src/Percolation.cpp                       */
  RVN_OP_SOILEVAP_ALL,         /* CmvSoilEvap             src/SoilEvaporation.cpp:379-383,767-783  */
  RVN_OP_SOILEVAP_HBV,         /* CmvSoilEvap             src/SoilEvaporation.cpp:385-...          */
  RVN_OP_SOILEVAP_GR4J,        /* CmvSoilEvap             src/SoilEvaporation.cpp                   */
  RVN_OP_CONVOLVE,             /* CmvConvolution          src/Convolution.cpp:245-303              */
  RVN_OP_CANEVP_ALL,           /* CmvCanopyEvap           src/VegetationMovers.cpp                  */
  RVN_OP_CANSNOWEVP_ALL,       /* CmvCanopySublimation    src/VegetationMovers.cpp                  */
  RVN_OP_OPEN_WATER_EVAP,      /* CmvOWEvaporation        src/OpenWaterEvap.cpp                     */
  RVN_OP_LAKE_EVAP_BASIC,      /* CmvLakeEvaporation      src/OpenWaterEvap.cpp                     */
  RVN_OP_CAPRISE_HBV,          /* CmvCapillaryRise        src/CapillaryRise.cpp:103-147,160-179    */
  RVN_OP_GLACIERMELT_HBV,      /* CmvGlacierMelt          src/GlacierProcesses.cpp                  */
  RVN_OP_GLACIERRELEASE_HBVEC, /* CmvGlacierRelease       src/GlacierProcesses.cpp                  */
  RVN_OP_INF_VIC_ARNO,         /* CmvInfiltration         src/Infiltration.cpp:400-416              */
  RVN_OP_BASE_LINEAR_ANALYTIC, /* CmvBaseflow             src/Baseflow.cpp:199-205                  */
  RVN_OP_BASE_VIC,             /* CmvBaseflow             src/Baseflow.cpp:221-228                  */
  RVN_OP_BASE_TOPMODEL,        /* CmvBaseflow             src/Baseflow.cpp:230-242                  */
  RVN_OP_SOILEVAP_TOPMODEL,    /* CmvSoilEvap             src/SoilEvaporation.cpp:385-402           */
  RVN_OP_INTERFLOW_PRMS,       /* CmvInterflow            src/Interflow.cpp:98-124,136-148          */
  RVN_OP_SNOWMELT_POTMELT,     /* CmvSnowMelt             src/SnowMeltRefreeze.cpp:87-131           */
  RVN_OP_SNOWSQUEEZE,          /* CmvSnowSqueeze          src/SnowMeltRefreeze.cpp:197-244          */
  RVN_OP_BASE_LINEAR_CONSTRAIN,/* CmvBaseflow             src/Baseflow.cpp:188-197                  */
  RVN_OP_BASE_CONSTANT,        /* CmvBaseflow             src/Baseflow.cpp:207-210                  */
  RVN_OP_BASE_THRESH_POWER,    /* CmvBaseflow             src/Baseflow.cpp:250-266                  */
  RVN_OP_BASE_THRESH_STOR,     /* CmvBaseflow             src/Baseflow.cpp:268-279                  */
  RVN_OP_PERC_GAWSER,          /* CmvPercolation          src/Percolation.cpp:204-213               */
  RVN_OP_PERC_GAWSER_CONSTRAIN,/* CmvPercolation          src/Percolation.cpp:215-226               */
  RVN_OP_PERC_POWER_LAW,       /* CmvPercolation          src/Percolation.cpp:228-236               */
  RVN_OP_PERC_LINEAR_ANALYTIC, /* CmvPercolation          src/Percolation.cpp:245-251               */
  RVN_OP_PERC_PRMS,            /* CmvPercolation          src/Percolation.cpp:253-267               */
  RVN_OP_PERC_SACRAMENTO,      /* CmvPercolation          src/Percolation.cpp:269-294               */
  RVN_OP_PERC_ASPEN,           /* CmvPercolation          src/Percolation.cpp:317-321               */
  RVN_OP_INF_RATIONAL,         /* CmvInfiltration         src/Infiltration.cpp:333-338              */
  RVN_OP_INF_ALL_INFILTRATES,  /* CmvInfiltration         src/Infiltration.cpp:347-353              */
  RVN_OP_INF_VIC,              /* CmvInfiltration         src/Infiltration.cpp:361-383              */
  RVN_OP_INF_PRMS,             /* CmvInfiltration         src/Infiltration.cpp:418-431              */
  RVN_OP_INF_PDM,              /* CmvInfiltration         src/Infiltration.cpp:540-563              */
  RVN_OP_SOILEVAP_LINEAR,      /* CmvSoilEvap             src/SoilEvaporation.cpp:370-377           */
  RVN_OP_SOILEVAP_PDM,         /* CmvSoilEvap             src/SoilEvaporation.cpp:424-437           */
  RVN_OP_SOILEVAP_HYMOD2,      /* CmvSoilEvap             src/SoilEvaporation.cpp:439-458           */
  RVN_OP_SOILEVAP_UBC,         /* CmvSoilEvap             src/SoilEvaporation.cpp:470-494           */
  RVN_OP_SOILEVAP_VIC,         /* CmvSoilEvap             src/SoilEvaporation.cpp:496-515           */
  RVN_OP_SOILEVAP_HYPR,        /* CmvSoilEvap             src/SoilEvaporation.cpp:385-420           */
  RVN_OP_SOILEVAP_SEQUEN,      /* CmvSoilEvap             src/SoilEvaporation.cpp:553-568           */
  RVN_OP_SOILEVAP_ROOT,        /* CmvSoilEvap             src/SoilEvaporation.cpp:570-596           */
  RVN_OP_SOILEVAP_ROOT_CONSTRAIN, /* CmvSoilEvap          src/SoilEvaporation.cpp:598-622           */
  RVN_OP_SOILEVAP_AWBM,        /* CmvSoilEvap             src/SoilEvaporation.cpp:716-727           */
  RVN_OP_SUBLIMATION,          /* CmvSublimation          src/Sublimation.cpp:98-243; ia[0] = rvn_sublim */
  RVN_OP_ABSTRACTION,          /* CmvAbstraction          src/Abstraction.cpp:230-256,514-532; ia[0] = rvn_abstraction */
  RVN_OP_CANEVP_MAXIMUM,       /* CmvCanopyEvap           src/VegetationMovers.cpp:141-145                  */
  RVN_OP_CANEVP_RUTTER,        /* CmvCanopyEvap           src/VegetationMovers.cpp:133-140 (no TRUNK variable) */
  RVN_OP_CANSNOWEVP_MAXIMUM,   /* CmvCanopySublimation    src/VegetationMovers.cpp:312-317                  */
  RVN_OP_INF_GREEN_AMPT,       /* CmvInfiltration         src/GreenAmpt.cpp:26-52,65-175; LambertN src/CommonFunctions.cpp:1572-1579 */
  RVN_OP_INF_GA_SIMPLE,        /* CmvInfiltration         src/GreenAmpt.cpp:65-175 (event memory in CUM_INFIL / GA_MOISTURE_INIT, connections src/Infiltration.cpp:25-33) */
  RVN_OP_INF_UBC,              /* CmvInfiltration         src/Infiltration.cpp:685-759 (GetUBCWMRunoff; connections :34-43; needs 4 soil layers) */
  RVN_OP_SNOBAL_COLD_CONTENT,  /* CmvSnowBalance          src/SnowBalance.cpp:732-835 (ColdContentBalance, Brook90 SNOENRGY / SNOPACK; connections :41-55); needs day_length */
  RVN_OP_SNOBAL_GAWSER,        /* CmvSnowBalance          src/SnowBalance.cpp:1098-1201 (GawserBalance: refreeze, density-based compaction, fixed sublimation; connections :123-145) */
  RVN_OP_SNOBAL_CRHM_EBSM,     /* CmvSnowBalance          src/SnowBalance.cpp:1214-1291 (CRHMSnowBalance, Marks et al. energy budget; connections :146-158) */
  RVN_OP_INF_SCS,              /* CmvInfiltration         src/Infiltration.cpp:338-343,638-680 (GetSCSRunoff: curve number with antecedent-moisture classes from the 5-day precipitation) */
  RVN_OP_INF_SCS_NOABSTRACTION,/* CmvInfiltration         same, Ia = 0 (src/Infiltration.cpp:668-677) */
  RVN_OP_SOILEVAP_SACSMA,      /* CmvSoilEvap             src/SoilEvaporation.cpp:652-715 (Sacramento soil accounting: six stores; connections :86-97) */
  RVN_OP_SOILEVAP_CHU,         /* CmvSoilEvap             src/SoilEvaporation.cpp:460-468 (Ontario crop-heat-unit method) */
  RVN_OP_CHU_ONTARIO,          /* CmvCropHeatUnitEvolve   src/CropGrowth.cpp:91-136 (CROP_HEAT_UNITS self-connection) */
  RVN_OP_EXCHANGE_FLOW,        /* CmvExchangeFlow         src/Flush.cpp:296-308,358-368: two soil stores swap EXCHANGE_FLOW [mm/d] in both directions; ia[0] = soil layer of the 'from' store */
  RVN_OP_GLACIERMELT_SIMPLE,   /* CmvGlacierMelt          GMELT_SIMPLE_MELT src/GlacierProcesses.cpp:126-129 */
  RVN_OP_GLACIERMELT_UBC,      /* CmvGlacierMelt          GMELT_UBC :139-166 (GLACIER_ICE -> PONDED_WATER, GLACIER_CC self-connection :29-36) */
  RVN_OP_GLACIERRELEASE_LINEAR,/* CmvGlacierRelease       GRELEASE_LINEAR / GRELEASE_LINEAR_ANALYTIC :317-327; ia[0] = 1 for the analytic form */
  RVN_OP_DEPRESSION_OVERFLOW,  /* CmvDepressionOverflow   src/DepressionProcesses.cpp:106-165; ia[0] = 0 DFLOW_THRESHPOW, 1 DFLOW_LINEAR, 2 DFLOW_WEIR */
  RVN_OP_SEEPAGE,              /* CmvSeepage SEEP_LINEAR  src/DepressionProcesses.cpp:254-298 (DEPRESSION -> soil / groundwater store) */
  RVN_OP_OPEN_WATER_RIPARIAN,  /* CmvOWEvaporation        OPEN_WATER_RIPARIAN src/OpenWaterEvap.cpp:137-141 (x STREAM_FRACTION) */
  RVN_OP_COUNT
} rvn_opcode;

/* model-wide method switches; numbering is local to this ABI (the host maps Raven's keywords) */
typedef enum { RVN_RAINSNOW_DATA = 0, RVN_RAINSNOW_DINGMAN, RVN_RAINSNOW_HBV, RVN_RAINSNOW_UBCWM, RVN_RAINSNOW_THRESHOLD } rvn_rainsnow;
typedef enum { RVN_POTMELT_NONE = 0, RVN_POTMELT_DATA, RVN_POTMELT_DEGREE_DAY, RVN_POTMELT_HBV, RVN_POTMELT_HMETS,
               RVN_POTMELT_DD_FREEZE, RVN_POTMELT_HBV_ROS, RVN_POTMELT_EB, RVN_POTMELT_RESTRICTED } rvn_potmelt;   /* src/PotentialMelt.cpp:20-246 */
/* atmospheric estimators of the forcing chain (reference src/UpdateForcings.cpp:677-731, src/Radiation.cpp:24-57,333-364) */
typedef enum { RVN_AIRPRESS_BASIC = 0, RVN_AIRPRESS_UBC, RVN_AIRPRESS_CONST } rvn_airpress;
typedef enum { RVN_RELHUM_CONSTANT = 0, RVN_RELHUM_MINDEWPT } rvn_relhum;
typedef enum { RVN_ABST_PERCENTAGE = 0, RVN_ABST_FILL, RVN_ABST_SCS, RVN_ABST_PDMROF } rvn_abstraction;   /* src/Abstraction.cpp:243-309 */
/* :PrecipIceptFract (reference precip_icept_method; CVegetationClass::RecalculateCanopyParams, src/VegetationParams.cpp:143-166) */
typedef enum { RVN_ICEPT_USER = 0, RVN_ICEPT_LAI, RVN_ICEPT_EXPLAI, RVN_ICEPT_NONE } rvn_icept;
typedef enum { RVN_WINDVEL_CONSTANT = 0, RVN_WINDVEL_UBC_MOD, RVN_WINDVEL_SQRT } rvn_windvel;   /* CModel::EstimateWindVelocity, src/UpdateForcings.cpp:739-790 */
/* snow sublimation algorithms (reference sublimation_type, src/Sublimation.cpp:98-200); SUBLIM_PBSM is a stub there */
typedef enum { RVN_SUBLIM_SVERDRUP = 0, RVN_SUBLIM_KUZMIN, RVN_SUBLIM_KUCHMENT_GELFAN, RVN_SUBLIM_BULK_AERO, RVN_SUBLIM_CENTRAL_SIERRA } rvn_sublim;
typedef enum { RVN_SW_RAD_DEFAULT = 0, RVN_SW_RAD_NONE } rvn_sw_rad;
typedef enum { RVN_SW_CLOUD_CORR_NONE = 0, RVN_SW_CLOUD_CORR_DINGMAN, RVN_SW_CLOUD_CORR_ANNANDALE } rvn_sw_cloudcorr;
typedef enum { RVN_PET_DATA = 0, RVN_PET_FROMMONTHLY, RVN_PET_CONSTANT, RVN_PET_NONE, RVN_PET_HARGREAVES_1985, RVN_PET_OUDIN,
               RVN_PET_HAMON, RVN_PET_LINEAR_TEMP, RVN_PET_MOHYSE,
               RVN_PET_TURC_1961, RVN_PET_MAKKINK_1957, RVN_PET_PENMAN_SIMPLE33, RVN_PET_PENMAN_SIMPLE39, RVN_PET_VAPDEFICIT, RVN_PET_LINACRE,
               RVN_PET_HARGREAVES, RVN_PET_PRIESTLEY_TAYLOR, RVN_PET_PENMAN_COMBINATION } rvn_pet;   /* CModel::EstimatePET, src/Evaporation.cpp:282-540 */
typedef enum { RVN_OROCORR_NONE = 0, RVN_OROCORR_SIMPLELAPSE, RVN_OROCORR_HBV } rvn_orocorr;
typedef enum { RVN_ROUTE_NONE = 0, RVN_ROUTE_MUSKINGUM, RVN_ROUTE_MUSKINGUM_CUNGE, RVN_ROUTE_STORAGECOEFF,
               RVN_ROUTE_PLUG_FLOW, RVN_ROUTE_DIFFUSIVE_WAVE, RVN_ROUTE_HYDROLOGIC, RVN_ROUTE_DIFFUSIVE_VARY } rvn_routing;
/* forcing perturbations of the EnKF members (reference forcing_type / adjustment, src/RavenInclude.h; applied by
 * CModel::ApplyForcingPerturbation, src/Model.cpp:2856-2879, in the order the reference calls it from
 * UpdateHRUForcingFunctions: TEMP_AVE after the temperature lapse, then PRECIP, RAINFALL, SNOWFALL after the precipitation lapse) */
typedef enum { RVN_PF_PRECIP = 0, RVN_PF_RAINFALL, RVN_PF_SNOWFALL, RVN_PF_TEMP_AVE } rvn_pert_forcing;
typedef enum { RVN_ADJ_ADDITIVE = 0, RVN_ADJ_MULTIPLICATIVE, RVN_ADJ_REPLACE } rvn_adjustment;
#define RVN_MAX_PERT 8
typedef enum { RVN_SOL_ORDERED_SERIES = 0, RVN_SOL_EULER = 1, RVN_SOL_ITERATED_HEUN = 2 } rvn_sol_method;
typedef enum { RVN_CONVOL_GR4J_1 = 0, RVN_CONVOL_GR4J_2, RVN_CONVOL_GAMMA, RVN_CONVOL_GAMMA_2 } rvn_convol_type;

/* ---- per-member parameter table ------------------------------------------------------------------
 * ptab[member][..] = [ globals | land-use classes | vegetation classes | soil classes | terrain classes ], every class
 * a fixed-stride row of the fields below (names follow reference src/Properties.h).  Members of an
 * ensemble differ only in this table (reference CModel::UpdateParameter, src/Model.cpp:2692-2750). */
typedef enum {
  RVN_G_SNOW_SWI = 0, RVN_G_SNOW_SWI_MIN, RVN_G_SNOW_SWI_MAX, RVN_G_SWI_REDUCT_COEFF, RVN_G_ADIABATIC_LAPSE,
  RVN_G_PRECIP_LAPSE, RVN_G_RAINSNOW_TEMP, RVN_G_RAINSNOW_DELTA, RVN_G_AVG_ANNUAL_SNOW, RVN_G_SNOW_TEMPERATURE,
  RVN_G_HBVEC_LAPSE_UPPER, RVN_G_HBVEC_LAPSE_RATE, RVN_G_HBVEC_LAPSE_ELEV, RVN_G_AIRSNOW_COEFF,
  RVN_G_AVG_ANNUAL_RUNOFF, RVN_G_MAX_SWE_SURFACE, RVN_G_TOC_MULTIPLIER, RVN_G_TIME_TO_PEAK_MULTIPLIER,
  RVN_G_GAMMA_SHAPE_MULTIPLIER, RVN_G_GAMMA_SCALE_MULTIPLIER, RVN_G_MOHYSE_PET_COEFF,
  RVN_G_SNOW_ROUGHNESS, RVN_G_WINDVEL_ICEPT, RVN_G_WINDVEL_SCALE, RVN_G_UBC_GW_SPLIT, RVN_G_UBC_FLASH_PONDING, RVN_G_COUNT
} rvn_global_field;
typedef enum {
  RVN_LU_IMPERMEABLE_FRAC = 0, RVN_LU_FOREST_COVERAGE, RVN_LU_FOREST_SPARSENESS, RVN_LU_MELT_FACTOR,
  RVN_LU_MIN_MELT_FACTOR, RVN_LU_MAX_MELT_FACTOR, RVN_LU_DD_MELT_TEMP, RVN_LU_DD_AGGRADATION,
  RVN_LU_REFREEZE_FACTOR, RVN_LU_DD_REFREEZE_TEMP, RVN_LU_REFREEZE_EXP, RVN_LU_HMETS_RUNOFF_COEFF,
  RVN_LU_GAMMA_SHAPE, RVN_LU_GAMMA_SCALE, RVN_LU_GAMMA_SHAPE2, RVN_LU_GAMMA_SCALE2, RVN_LU_GR4J_X4,
  RVN_LU_HBV_MELT_FOR_CORR, RVN_LU_HBV_MELT_ASP_CORR, RVN_LU_HBV_MELT_GLACIER_CORR, RVN_LU_HBV_GLACIER_KMIN,
  RVN_LU_GLAC_STORAGE_COEFF, RVN_LU_HBV_GLACIER_AG, RVN_LU_DEP_MAX, RVN_LU_LAKE_PET_CORR, RVN_LU_OW_PET_CORR,
  RVN_LU_PARTITION_COEFF, RVN_LU_CC_DECAY_COEFF, RVN_LU_MAX_SAT_AREA_FRAC, RVN_LU_PDM_B, RVN_LU_AET_COEFF,
  RVN_LU_MAX_DEP_AREA_FRAC, RVN_LU_PONDED_EXP, RVN_LU_AWBM_AREAFRAC1, RVN_LU_AWBM_AREAFRAC2, RVN_LU_HYMOD2_G, RVN_LU_HYMOD2_KMAX,
  RVN_LU_HYMOD2_EXP, RVN_LU_PET_LIN_COEFF, RVN_LU_RAIN_MELT_MULT, RVN_LU_RELHUM_CORR, RVN_LU_PET_VAP_COEFF,
  RVN_LU_MIN_WIND_SPEED, RVN_LU_MAX_WIND_SPEED, RVN_LU_ABST_PERCENT, RVN_LU_ROUGHNESS, RVN_LU_PRIESTLEYTAYLOR_COEFF, RVN_LU_SCS_CN, RVN_LU_SCS_IA_FRACTION, RVN_LU_PDMROF_B,
  RVN_LU_DEP_THRESHOLD, RVN_LU_DEP_N, RVN_LU_DEP_MAX_FLOW, RVN_LU_DEP_K, RVN_LU_DEP_CRESTRATIO, RVN_LU_DEP_SEEP_K, RVN_LU_STREAM_FRACTION, RVN_LU_COUNT
} rvn_landuse_field;
typedef enum {
  RVN_V_MAX_HEIGHT = 0, RVN_V_MAX_LAI, RVN_V_SAI_HT_RATIO, RVN_V_MAX_CAPACITY, RVN_V_MAX_SNOW_CAPACITY,
  RVN_V_RAIN_ICEPT_PCT, RVN_V_SNOW_ICEPT_PCT, RVN_V_PET_VEG_CORR, RVN_V_RAIN_ICEPT_FACT, RVN_V_SNOW_ICEPT_FACT, RVN_V_ALBEDO, RVN_V_SVF_EXTINCTION, RVN_V_CHU_MATURITY,
  /* derived canopy variables (reference veg_var_struct; CVegetationClass::RecalculateCanopyParams,
   * src/VegetationParams.cpp:85-200), refreshed by the host for the current step when date-dependent */
  RVN_V_VAR_CAPACITY, RVN_V_VAR_SNOW_CAPACITY, RVN_V_VAR_RAIN_ICEPT_PCT, RVN_V_VAR_SNOW_ICEPT_PCT,
  RVN_V_COUNT
} rvn_veg_field;
typedef enum {
  RVN_S_POROSITY = 0, RVN_S_CAP_RATIO, RVN_S_PERC_COEFF, RVN_S_PET_CORRECTION, RVN_S_BASEFLOW_COEFF,
  RVN_S_BASEFLOW_N, RVN_S_FIELD_CAPACITY, RVN_S_SAT_WILT, RVN_S_HBV_BETA, RVN_S_MAX_CAP_RISE_RATE,
  RVN_S_MAX_PERC_RATE, RVN_S_GR4J_X2, RVN_S_GR4J_X3, RVN_S_VIC_B_EXP, RVN_S_MAX_BASEFLOW_RATE, RVN_S_MAX_INTERFLOW_RATE,
  RVN_S_PERC_N, RVN_S_SAC_PERC_ALPHA, RVN_S_SAC_PERC_EXPON, RVN_S_PERC_ASPEN, RVN_S_BASEFLOW_THRESH, RVN_S_BASEFLOW_COEFF2,
  RVN_S_STORAGE_THRESHOLD, RVN_S_VIC_ZMIN, RVN_S_VIC_ZMAX, RVN_S_VIC_ALPHA, RVN_S_VIC_EVAP_GAMMA, RVN_S_UBC_EVAP_SOIL_DEF,
  RVN_S_UBC_INFIL_SOIL_DEF, RVN_S_ALBEDO_WET, RVN_S_ALBEDO_DRY, RVN_S_HYDRAUL_COND, RVN_S_WETTING_FRONT_PSI, RVN_S_UNAVAIL_FRAC, RVN_S_EXCHANGE_FLOW, RVN_S_COUNT
} rvn_soil_field;
typedef enum { RVN_T_TOPMODEL_LAMBDA = 0, RVN_T_COUNT } rvn_terrain_field;   /* reference terrain_struct, src/Properties.h */

/* forcing series the host uploads per gauge, already sampled on the model step
 * (reference CGauge::GetForcingValue(ftype,nn), src/Gauge.cpp:533-540) */
typedef enum {
  RVN_GF_PRECIP = 0, RVN_GF_SNOW_FRAC, RVN_GF_TEMP_AVE, RVN_GF_TEMP_DAILY_AVE, RVN_GF_TEMP_DAILY_MIN,
  RVN_GF_TEMP_DAILY_MAX, RVN_GF_PET, RVN_GF_OW_PET, RVN_GF_POTENTIAL_MELT, RVN_GF_COUNT
} rvn_gauge_series;
/* per-gauge constants */
typedef enum {
  RVN_GC_ELEVATION = 0, RVN_GC_RAIN_CORR, RVN_GC_SNOW_CORR, RVN_GC_TEMP_CORR,
  RVN_GC_MONTHLY_AVE_TEMP0 /*..+11*/ = 4, RVN_GC_MONTHLY_AVE_PET0 /*..+11*/ = 16,
  RVN_GC_MONTHLY_MIN_TEMP0 = 28, RVN_GC_MONTHLY_MAX_TEMP0 = 40, RVN_GC_COUNT = 52
} rvn_gauge_const;

/* derived per-HRU forcings kept on the device (subset of reference force_struct, src/RavenInclude.h:1294-1343) */
typedef enum {
  RVN_F_PRECIP = 0, RVN_F_SNOW_FRAC, RVN_F_TEMP_AVE, RVN_F_TEMP_DAILY_AVE, RVN_F_TEMP_DAILY_MIN,
  RVN_F_TEMP_DAILY_MAX, RVN_F_PET, RVN_F_OW_PET, RVN_F_POTENTIAL_MELT, RVN_F_TEMP_MONTH_AVE, RVN_F_PET_MONTH_AVE,
  /* atmospheric forcings: derived only when opt.need_atmos, recorded (rvn_gpu_download_forcings) but not carried through the sweep */
  RVN_F_AIR_PRES, RVN_F_REL_HUMIDITY, RVN_F_SW_RADIA, RVN_F_WIND_VEL,
  /* net radiation (opt.need_radiation): shortwave above the canopy x (1 - total albedo), net longwave */
  RVN_F_SW_RADIA_NET, RVN_F_LW_RADIA_NET,
  RVN_F_COUNT
} rvn_forcing_field;
#define RVN_F_CARRIED RVN_F_AIR_PRES   /* fields below this index live in the per-thread forcing vector of the sweep */

/* per-step calendar record (reference time_struct, filled by JulianConvert, src/CommonFunctions.cpp:300-380) */
typedef struct rvn_time {
  double model_time;   /* [d] since start */
  double julian_day;   /* [d] fractional day of year, 0 = Jan 1 00:00 */
  double day_angle;    /* CRadiation::DayAngle(mid_day,year) [rad] */
  int32_t month;       /* 1..12 */
  int32_t day_of_month;
  int32_t year;
  int32_t day_changed; /* tt.day_changed */
  int32_t nn;          /* sample index into the gauge series */
  int32_t pad_;
} rvn_time;

typedef struct rvn_process {
  int32_t op;       /* rvn_opcode */
  int32_t nconn;    /* reference GetNumConnections() (101 for a convolution) */
  int32_t conn0;    /* offset of this process' first connection in conn_from/conn_to (convolutions store none) */
  int32_t ia[5];    /* op-specific ints:
                       CONVOLVE: ia[0]=conv index, ia[1]=target SV, ia[2]=first CONV_STOR SV, ia[3]=CONVOLUTION SV,
                                 ia[4]=rvn_convol_type
                       soil ops: ia[0]=soil layer of 'from', ia[1]=soil layer of 'to' (or -1)           */
  double  da[2];    /* op-specific doubles (SPLIT: da[0] = fraction to first target) */
  /* :ProcessGroup ... :EndProcessGroup (reference CProcessGroup, src/ProcessGroup.cpp:72-99): the members of a group are
   * consecutive records; every member's rates are evaluated on the SAME state into a shared local rate vector that is zeroed
   * once per group (the reference's loc_rates), scaled by `weight` into the group's rate vector, and the group's connections
   * (the concatenation of the members', conn_from/conn_to[gconn0 .. gconn0+gnconn)) are applied after the last member. */
  int32_t group;    /* 0: ordinary process; bit 0: first member of a group; bit 1: last member; 4: middle member */
  int32_t gconn0, gnconn, pad_;
  double  weight;
} rvn_process;
#define RVN_GRP_FIRST 1
#define RVN_GRP_LAST 2
#define RVN_GRP_MEMBER 4

/* one lateral exchange process (LAT_FLUSH: reference CmvLatFlush, src/LatFlush.cpp:55-234) */
typedef enum { RVN_LAT_FLUSH = 0, RVN_LAT_EQUILIBRATE = 1 /* CmvLatEquilibrate, src/LatEquilibrate.cpp:49-212 */,
               RVN_LAT_REDISTRIBUTE = 2 /* CmvLatRedistribute, src/SnowRedistribute.cpp:59-215 */ } rvn_lateral_kind;
typedef enum { RVN_REDIST_CONTINUOUS = 0, RVN_REDIST_THRESHOLD = 1 } rvn_redist_method;
typedef struct rvn_lateral {
  int32_t kind;            /* rvn_lateral_kind */
  int32_t sv_from, sv_to;  /* state-variable indices _iFromLat / _iToLat (the same for every connection of a flush) */
  int32_t nconn, conn0;    /* connections in lat_kfrom/lat_kto/lat_frac */
  int32_t nchunks, chunk0; /* chunk offsets in lat_chunk_ptr (nchunks + 1 entries) */
  int32_t method;          /* LAT_REDISTRIBUTE: rvn_redist_method */
  double  mixing_rate;     /* LAT_EQUILIBRATE: fraction mixed per day, the step mixes min(mixing_rate*dt, 1);
                            * LAT_REDISTRIBUTE: _max_snow_height [mm] */
} rvn_lateral;

typedef struct rvn_options {
  double  timestep;               /* Options.timestep [d] */
  int32_t rainsnow, pot_melt, evaporation, ow_evaporation;
  int32_t orocorr_temp, orocorr_precip, orocorr_pet;
  int32_t routing, suppress_competitive_ET, snow_suppress_PET, allow_soil_overfill, direct_evap;
  int32_t lake_sv;                /* CModel::GetLakeStorageIndex() */
  int32_t keep_flux_balance;      /* 1: keep CModel::_aCumulativeBal (src/Model.cpp:2428-2436) on the device: the cumulative amount moved through every process
                                   * connection of every HRU, [member][nConn][nHRU], read with rvn_gpu_download_flux_balance.  Interpreter sweep, ORDERED_SERIES,
                                   * process lists without :Convolve (whose 2N+1 internal connections per process are not tabulated in conn_from/conn_to).
                                   * 0 (default): not kept - the array is larger than the state and nothing on the hot path reads it */
  int32_t sol_method;             /* rvn_sol_method: ORDERED_SERIES (default), EULER or ITERATED_HEUN (reference src/Solvers.cpp:205-251 / 256-297 / 304-397) */
  int32_t need_atmos;             /* 1: a consumed PET or a sublimation process needs air pressure / relative humidity / wind / shortwave radiation -> derive them */
  int32_t air_pressure, rel_humidity, sw_radiation, sw_cloudcorr;   /* rvn_airpress, rvn_relhum, rvn_sw_rad, rvn_sw_cloudcorr */
  int32_t wind_velocity;          /* rvn_windvel */
  int32_t need_radiation;         /* 1: a consumed PET or potential-melt method needs net shortwave / longwave radiation (implies need_atmos) */
  int32_t lw_radiation;           /* 0 LW_RAD_DEFAULT (Dingman clear-sky emissivity with LW_INC_DEFAULT, src/Radiation.cpp:210-231), 1 LW_RAD_NONE */
  int32_t sw_canopycorr;          /* 0 SW_CANOPY_CORR_NONE, 1 SW_CANOPY_CORR_STATIC (src/Radiation.cpp:377-415) */
  int32_t pad_opt_;
  int32_t interception_factor;    /* rvn_icept */
  int32_t max_iterations;         /* ITERATED_HEUN: Options.max_iterations (30) */
  double  convergence_crit;       /* ITERATED_HEUN: Options.convergence_crit (0.01) */
} rvn_options;

/* Flattened model (all pointers are HOST memory, copied by rvn_gpu_create). */
typedef struct rvn_model_desc {
  int32_t abi_version;
  int32_t nHRU, nSB, nMembers, nGauges, nSteps;
  rvn_options opt;

  /* state-variable catalogue, order == reference CModel::AddStateVariables (src/Model.cpp:1561-1612) */
  int32_t NS, nSoilLayers;
  const int32_t *sv_type;     /* [NS] */
  const int32_t *sv_layer;    /* [NS] */

  /* process program (reference CModel::_pProcesses order) */
  int32_t nProc, nConn;
  const rvn_process *proc;    /* [nProc] */
  const int32_t *conn_from;   /* [nConn] */
  const int32_t *conn_to;     /* [nConn] */
  const uint64_t *apply_mask; /* [nHRU] bit j = _aShouldApplyProcess[j][k] (src/ModelInitialize.cpp:252-263) */

  /* HRU statics (shared by all members) */
  const double *hru_area, *hru_elev, *hru_lat_rad, *hru_slope, *hru_aspect;   /* [nHRU] */
  const int32_t *hru_type, *hru_lu, *hru_veg, *hru_profile, *hru_sb, *hru_enabled; /* [nHRU] */

  /* soil profiles: layer m of profile r has class prof_class[r*RVN_MAX_SOIL_LAYERS+m], thickness [m] likewise */
  int32_t nProfiles, nLU, nVeg, nSoilClasses;
  const int32_t *prof_class;
  const double  *prof_thick;

  /* parameter tables [nMembers][ptab_stride]; offsets of the class blocks inside one member's table */
  int32_t ptab_stride, glob_off, lu_off, veg_off, soil_off;
  const double *ptab;
  /* terrain classes: block of nTerrain rows of RVN_T_COUNT fields at terr_off in every member's table; class of each HRU
   * (-1 = [NONE]) */
  int32_t terr_off, nTerrain;
  const int32_t *hru_terrain; /* [nHRU] */
  /* convolution unit hydrographs, hoisted out of the per-HRU-per-step loop (reference rebuilds them in
   * CmvConvolution::GenerateUnitHydrograph, src/Convolution.cpp:107-154): [nMembers][nLU][nConvProc] */
  int32_t nConvProc;
  const int32_t *conv_n;      /* N of each UH */
  const double  *conv_uh;     /* [..][RVN_CONV_STORES] */

  /* forcings */
  const double *gauge_series; /* [RVN_GF_COUNT][nGauges][nSteps] */
  const double *gauge_const;  /* [nGauges][RVN_GC_COUNT] */
  /* gauge (or grid-cell) interpolation weights as CSR over HRUs, entries in ascending gauge index: entry i of HRU k is
   * gauge wt_idx[i] with weights wt_val[3*i+{0,1,2}] = {precipitation, temperature, everything else} -- the three dense
   * [nHRU][nGauges] weight tables of the reference (src/Model.h:104-106, used at src/UpdateForcings.cpp:184-235), or the
   * sparse grid weights of CForcingGrid::GetWeightedValue (src/ForcingGrid.cpp:1886-1946).  Zero weights are omitted:
   * they contribute exact zeros to the reference's sums. */
  const int32_t *wt_ptr;      /* [nHRU+1] */
  const int32_t *wt_idx;      /* [nnz] */
  const double  *wt_val;      /* [nnz][3] */
  const rvn_time *times;      /* [nSteps] */

  /* routing network (reference CSubBasin + CModel::InitializeRoutingNetwork, src/ModelInitialize.cpp:505-619) */
  const int32_t *sb_down;     /* [nSB] downstream index or -1 */
  const int32_t *sb_order;    /* [nSB] _aOrderedSBind: processing order, upstream first */
  const int32_t *sb_level;    /* [nSB] _aSubBasinOrder: hops to outlet */
  const int32_t *sb_headwater;/* [nSB] */
  const int32_t *sb_nseg, *sb_nqin, *sb_nqlat; /* [nSB] */
  const double *sb_area;      /* [nSB] km2 */
  const double *sb_rain_corr, *sb_snow_corr, *sb_temp_corr; /* [nSB] */
  const double *sb_musk_K, *sb_musk_X;        /* [nSB] per-segment Muskingum K [d], X (src/SubBasin.cpp:2429-2452) */
  const double *sb_K_reach;   /* [nSB] GetMuskingumK(dx) for STORAGECOEFF */
  int32_t max_nqin, max_nqlat;
  const double *sb_unit_hydro;   /* [nSB][max_nqlat] _aUnitHydro */
  const double *sb_route_hydro;  /* [nSB][max_nqin]  _aRouteHydro */
  double watershed_area;         /* km2 */
  /* ROUTE_HYDROLOGIC (reference src/SubBasin.cpp:2685-2741): rating curves of the channel profiles, flow -> cross-section area
   * (CChannelXSect::_aQ/_aXArea; GetArea = InterpolateCurve(Q/Q_mult,...), src/ChannelXSect.cpp:297-302), and per basin the
   * profile index, the Manning/slope flow multiplier (GetFlowCorrections, :345-359) and the reach length [m].  NULL otherwise. */
  int32_t nChannels, nChanPts;
  const double *chan_Q, *chan_A;   /* [nChannels][nChanPts] */
  const int32_t *sb_channel;       /* [nSB] */
  const double *sb_Qmult, *sb_reach_m; /* [nSB] */
  /* ROUTE_DIFFUSIVE_VARY (reference CSubBasin::UpdateRoutingHydro, src/SubBasin.cpp:2521-2572): the routing hydrograph of every basin is
   * rebuilt each step from a history of celerities read off the channel rating curve at the latest inflow.  Uses chan_Q / chan_A /
   * sb_channel / sb_Qmult / sb_reach_m above plus, per basin, the reference celerity c_ref [m/s] and alpha = D_ref / 2 [m2/d] with D_ref
   * the diffusivity at the reference flow.  NULL otherwise. */
  const double *sb_c_ref, *sb_diff_alpha; /* [nSB] */
  /* month-interpolated relative vegetation height / LAI of every vegetation class at every step
   * (InterpolateMo(relative_ht|relative_LAI, tt), reference src/VegetationParams.cpp:102-107): [nSteps][nVeg][2] */
  const double *veg_season;
  /* extraterrestrial radiation on each HRU's slope [MJ/m2/d] (reference CRadiation::CalcETRadiation2, src/Radiation.cpp:
   * 637-822) for every distinct (day angle, step interval) of the run: et_radia[nEtRows][nHRU], row of step n = et_row[n].
   * NULL when no process consumes a PET estimated from it (PET_HARGREAVES_1985). */
  int32_t nEtRows, pad2_;
  const double *et_radia;
  const int32_t *et_row;      /* [nSteps] */
  /* day geometry on the same rows (NULL unless a consumed PET needs it): day_length[row][hru] = CRadiation::DayLength(lat, declin)
   * [d] (src/Radiation.cpp:438-445; F.day_length, src/UpdateForcings.cpp:175) for PET_HAMON, sunset_angle[row][hru] =
   * acos(-tan(lat)*tan(declin)) [rad] for PET_MOHYSE (src/Evaporation.cpp:476-483) */
  const double *day_length, *sunset_angle;
  /* clear-sky shortwave pieces on the same rows (NULL unless opt.need_atmos): et_flat = extraterrestrial radiation on a horizontal
   * surface (F.ET_radia_flat), opt_air_mass = CRadiation::OpticalAirMass (src/Radiation.cpp:909-967) */
  const double *et_flat, *opt_air_mass;
  /* lateral exchange processes (reference CLateralExchangeProcessABC; CModel::ApplyLateralProcess, src/Model.cpp:3126-3165,
   * applied by the solver after ALL vertical processes of ALL HRUs and before the routing prep, src/Solvers.cpp:404-443).
   * Only processes that apply (gate of the first source HRU, every participating HRU enabled) are listed, in process order.
   * Connection q of process L: lat_kfrom/lat_kto/lat_frac[L.conn0 + q].  Connections are grouped into chunks of HRUs that no
   * other chunk touches (one subbasin each; a single chunk for INTERBASIN): chunk c of L = connections
   * lat_chunk_ptr[L.chunk0 + c] .. lat_chunk_ptr[L.chunk0 + c + 1] (offsets relative to L.conn0). */
  int32_t nLat, nLatConn;
  const struct rvn_lateral *lat;   /* [nLat] */
  const int32_t *lat_kfrom, *lat_kto;
  const double  *lat_frac;         /* FLUSH: share of the source's water this recipient gets; EQUILIBRATE: A_to / A_sum of its subbasin;
                                    * REDISTRIBUTE: the connection weight of :LateralConnections */
  const int32_t *lat_flag;         /* EQUILIBRATE (the control flow of src/LatEquilibrate.cpp:190-207, which depends on the connection
                                    * list only): bit 0 = first add mix*max(S_to,0)*A_to to the subbasin's pool; bit 1 = step 1 (mix the
                                    * source into the pool), else step 2 (hand the pool back by area) */
  const int32_t *lat_chunk_ptr;
  /* forcing perturbations (EnKF members; 0 for every other model).  One sample per perturbation, member and DAY (the
   * reference draws nStepsPerDay samples at the start of a day, src/Model.cpp:2829-2847; steps here are daily or longer):
   * pert_eps[member][step][pert].  pert_mask[pert][hru] = 1 where the HRU belongs to the perturbation's HRU group
   * (NULL = every perturbation applies to every HRU). */
  /* user-specified basin inflows [m3/s] (reference :BasinInflowHydrograph / :BasinInflowHydrograph2, CSubBasin::GetSpecifiedInflow /
   * GetDownstreamInflow): for the nSpec basins that have one (spec_basin[], ascending index) and every step, spec_in = the value the
   * solver adds to the basin's inflow from upstream (at t + dt, src/Solvers.cpp:543), spec_down = the value it adds to the outflow of
   * the last reach segment (at t, :586), spec_cum = the basin's term of CModel::IncrementCumulInput [mm] (src/Model.cpp:2467-2469).
   * Layout [nSteps][nSpec]; nSpec = 0: none. */
  int32_t nSpec, pad4_;
  const int32_t *spec_basin;
  const double *spec_in, *spec_down, *spec_cum;
  /* reservoirs / lakes at subbasin outlets (reference CReservoir, src/Reservoir.cpp; level-pool routing CReservoir::RouteWater
   * :1532-1853 called from the solver after the reach is routed, src/Solvers.cpp:591-596).  Reservoir r sits in basin res_basin[r]
   * (ascending basin index) and is linked to HRU res_hru[r] (-1: none; a linked HRU's surface water becomes precipitation on the
   * reservoir, its open-water PET x LAKE_PET_CORR the reservoir evaporation, and its AET state variable receives CReservoir::GetAET,
   * src/Solvers.cpp:502-503,608-612).  Curves (the reference's _aStage/_aQ/_aQunder/_aArea/_aVolume, built by the host exactly like the
   * CReservoir constructors) are stored [nRes][RVN_RES_CURVES][nResPts] with res_np[r] valid points each.  res_min_stage = _min_stage
   * (returned for a reservoir that dried out).  res_extract[nSteps][nRes] = demand of the reservoir's :ReservoirExtraction series for the
   * step [m3/s] (CDemand::UpdateDemand, src/WaterDemands.cpp:204-216; at most one series per reservoir), 0 without one.
   * res_state0[nRes][RVN_RES_STATE] = initial {stage, stage_last, Qout, Qout_last, precip m3, MB losses m3, AET m3, 0} of every member.
   * Supported: no control structures, DZTR, seepage, stage / flow constraint series or stage assimilation. */
  int32_t nRes, nResPts;
  const int32_t *res_basin, *res_hru, *res_np;   /* [nRes] */
  const double *res_curves, *res_min_stage, *res_extract, *res_state0;
  /* direct-insertion streamflow assimilation (reference src/Assimilate.cpp, :AssimilateStreamflow; DA_ECCC).  Everything the reference's
   * PrepareAssimilation derives from the observations and the network alone is tabulated by the host per step, [nSteps][nSB]:
   * da_override = 1 where a gauged basin has a valid observation for the step (its outflow is overridden, CModel::AssimilationOverride
   * :67-118), da_obsQ / da_obsQ2 the observation at the end of this and the next step (-1.2345 = blank), and for the upstream propagation
   * of the correction (AssimilationBackPropagate :283-339) da_decay = exp(-distfact * _aDAlength[p]) and da_wt = the drainage-area weight
   * ECCCwt.  Per basin: sb_drain_area [km2] and sb_da_corr = exp(-distfact * reach length) of CSubBasin::AdjustAllFlows.
   * assimilation_fact = the ASSIMILATION_FACT global.  assimilate_flow = 0: none (pointers NULL).  Not with reservoirs. */
  int32_t assimilate_flow, pad5_;
  double assimilation_fact;
  const uint8_t *da_override;
  const double *da_obsQ, *da_obsQ2, *da_decay, *da_wt;
  const double *sb_drain_area, *sb_da_corr;
  int32_t nPert, pad3_;
  const int32_t *pert_forcing;  /* [nPert] rvn_pert_forcing */
  const int32_t *pert_adj;      /* [nPert] rvn_adjustment */
  const uint8_t *pert_mask;     /* [nPert][nHRU] or NULL */
  const double  *pert_eps;      /* [nMembers][nSteps][nPert] */
  /* 5-day antecedent precipitation of every gauge at every step, or NULL when no process reads force_struct::precip_5day:
   * CGauge::GetForcingValue(F_PRECIP, t - 5, 5) * 5 (src/UpdateForcings.cpp:107; blank value before the series starts).  The sweep applies the
   * precipitation weights, gauge / basin corrections and the orographic correction to it exactly as to the precipitation (:525, src/OrographicCorrections.cpp:229,244) */
  const double *gauge_precip5;  /* [nGauges][nSteps] */
  /* force_struct::subdaily_corr of every HRU for every distinct (date, time of day) row of et_row, or NULL (= 1: daily steps or SUBDAILY_NONE).
   * SUBDAILY_SIMPLE (CModel::CalculateSubDailyCorrection, src/UpdateForcings.cpp:964-982): the share of the day's half-sine between dawn and dusk
   * that falls into the step, from the HRU's day length; multiplies the daily PET estimators and the degree-day potential melt */
  const double *subdaily_corr;  /* [nEtRows][nHRU] */
} rvn_model_desc;

/* ---- EnKF state matrix (reference CEnKFEnsemble::AddToStateMatrix / UpdateFromStateMatrix, src/EnKF.cpp:602-747) ------
 * One record per row of the reference's _state_matrix[e][...], in the reference's row order. */
typedef enum {
  RVN_XROW_HRU_SV = 0,   /* index = HRU k, sub = state-variable index i        : pHRU->GetStateVarValue(i)   */
  RVN_XROW_QIN_HIST,     /* index = subbasin p, sub = n                          : _aQinHist[n]               */
  RVN_XROW_QOUT,         /* index = subbasin p, sub = segment                    : _aQout[seg]                */
  RVN_XROW_QOUT_LAST     /* index = subbasin p                                   : _QoutLast                  */
} rvn_xrow_kind;
typedef struct rvn_enkf_row { int32_t kind, index, sub, pad_; } rvn_enkf_row;

typedef struct rvn_gpu_model rvn_gpu_model; /* opaque */

/* ---- entry points ------------------------------------------------------------------------------- */
const char *rvn_gpu_last_error(void);
int rvn_gpu_abi_version(void);
/* sha256 of the sources + compiler flags this library was built from (ravenhydroframework_b200/build.py stamps it and rebuilds
 * on mismatch), "unstamped" for a hand build */
const char *rvn_gpu_build_digest(void);
int rvn_gpu_device_count(void);

/* build device-resident model; `device` = CUDA ordinal */
int rvn_gpu_create(const rvn_model_desc *desc, int device, rvn_gpu_model **out);
/* the same with RVN_CREATE_* flags.  RVN_CREATE_FORCE_INTERPRETER: serve a model whose process list matches a compiled-in
 * template program with the generic interpreter kernel anyway (both must give bit-identical results; used by the tests) */
#define RVN_CREATE_FORCE_INTERPRETER 1
int rvn_gpu_create_ex(const rvn_model_desc *desc, int device, int flags, rvn_gpu_model **out);
void rvn_gpu_destroy(rvn_gpu_model *m);

/* HRU state, host layout [member][hru][sv] (the reference's aPhi[k][i] per member); device is SoA */
int rvn_gpu_upload_state(rvn_gpu_model *m, const double *state, int member0, int nmembers);
int rvn_gpu_download_state(rvn_gpu_model *m, double *state, int member0, int nmembers);
/* basin routing state, host layout per member: qin_hist[nSB][max_nqin], qlat_hist[nSB][max_nqlat],
 * qout[nSB][RVN_MAX_RIVER_SEGS], scalars[nSB][8] = {QoutLast,QlatLast,Qlocal,QlocLast,channel_storage,rivulet_storage,0,0} */
int rvn_gpu_upload_basin_state(rvn_gpu_model *m, const double *qin_hist, const double *qlat_hist,
                               const double *qout, const double *scalars, int member0, int nmembers);
int rvn_gpu_download_basin_state(rvn_gpu_model *m, double *qin_hist, double *qlat_hist, double *qout,
                                 double *scalars, int member0, int nmembers);
/* reservoir state, host layout [member][nRes][RVN_RES_STATE] (see rvn_model_desc::res_state0); no-ops for a model without reservoirs */
int rvn_gpu_upload_reservoir_state(rvn_gpu_model *m, const double *res, int member0, int nmembers);
int rvn_gpu_download_reservoir_state(rvn_gpu_model *m, double *res, int member0, int nmembers);
/* replace parameter tables / convolution UHs of a member range (ensemble UpdateModel) */
int rvn_gpu_upload_params(rvn_gpu_model *m, const double *ptab, const int32_t *conv_n, const double *conv_uh,
                          int member0, int nmembers);
/* replace forcing series (streaming long runs / BMI SetValue): same layout as desc, nsteps samples from step0 */
/* CModel::GetCumulativeFlux / GetCumulFluxBetween source data (src/Model.cpp:902-965): cumulative [mm] moved through connection q (conn_from[q] -> conn_to[q],
 * process-list order; members of a :ProcessGroup: the group's weighted connections) of HRU k, for members member0 .. member0+nmembers-1: cum[member][nConn][nHRU].
 * Needs opt.keep_flux_balance = 1 at create. */
int rvn_gpu_download_flux_balance(rvn_gpu_model *m, double *cum, int member0, int nmembers);
int rvn_gpu_upload_forcings(rvn_gpu_model *m, const double *gauge_series, const rvn_time *times, int step0, int nsteps);

/* advance all members by nsteps model steps starting at step index `step0`:
 * forcing update -> HRU process sweep -> HRU->basin aggregation -> ordered routing -> balance reductions
 * (= UpdateHRUForcingFunctions + MassEnergyBalance + IncrementCumulInput/CumOutflow of the reference loop,
 * src/RavenMain.cpp:151-160).  Asynchronous on the handle's stream; rvn_gpu_sync waits. */
int rvn_gpu_run(rvn_gpu_model *m, int step0, int nsteps);
int rvn_gpu_sync(rvn_gpu_model *m);
/* streaming form of ONE model step for host-side callers (BMI-style drivers that feed forcings and consume hydrographs step by
 * step): nothing waits.  Enqueued, in order: copy of this step's forcing slab forcing_host[nGauges][RVN_GF_COUNT] (station-major: the layout the sweep gathers from; pinned host
 * memory, may be NULL to keep the uploaded series) to the device; the step; copies of the step's outflows
 * qout_host[nMembers][nSB] and watershed record wshed_host[nMembers][4] (pinned, either may be NULL) back to the host.  The
 * result copies run on the routing stream underneath the next step's sweep.  Host buffers must stay untouched / are valid
 * after rvn_gpu_sync (or any later synchronising call); steps must be enqueued in order. */
int rvn_gpu_step_async(rvn_gpu_model *m, int step, const double *forcing_host, double *qout_host, double *wshed_host);

/* per-step outputs recorded on the device during rvn_gpu_run (max_steps records; 0 disables):
 *   flags bit 0: qout  [rec][member][nSB]  CSubBasin::GetOutflowRate() after the step
 *   flags bit 1: wshed [rec][member][4]    {_CumulInput, _CumulOutput, avg precip, total water storage [mm]}
 *                (the WatershedStorage.csv columns, src/StandardOutput.cpp:745-785) */
int rvn_gpu_set_recording(rvn_gpu_model *m, int max_steps, int flags);
/* keep the derived per-HRU forcings of the last step on the device (off by default: 88 B/HRU/step) */
int rvn_gpu_enable_forcing_output(rvn_gpu_model *m, int on);
/* bracket every HRU-sweep launch with CUDA events so that rvn_gpu_last_run_stats can report sweep_ms */
int rvn_gpu_set_kernel_timing(rvn_gpu_model *m, int on);
/* {_CumulInput,_CumulOutput} [mm] per member: out[member][2] */
int rvn_gpu_download_cumulative(rvn_gpu_model *m, double *cumul, int member0, int nmembers);
int rvn_gpu_download_hydrographs(rvn_gpu_model *m, double *qout, int rec0, int nrec, int member0, int nmembers);
int rvn_gpu_download_watershed(rvn_gpu_model *m, double *wshed, int rec0, int nrec, int member0, int nmembers);
/* derived per-HRU forcings of the last executed step, host layout [member][hru][RVN_F_COUNT] */
int rvn_gpu_download_forcings(rvn_gpu_model *m, double *forc, int member0, int nmembers);

/* area-weighted watershed average of every state variable (reference CModel::GetAvgStateVar,
 * src/Model.cpp:1159-1174): out[member][NS] */
int rvn_gpu_avg_state(rvn_gpu_model *m, double *out, int member0, int nmembers);

/* replace the forcing-perturbation samples of a member range (next assimilation window): eps[nmembers][nSteps][nPert] */
int rvn_gpu_upload_perturbations(rvn_gpu_model *m, const double *eps, int member0, int nmembers);

/* ---- EnKF analysis on the device (reference CEnKFEnsemble::AssimilationCalcs, src/EnKF.cpp:478-596) -------------------
 * rvn_gpu_enkf_configure  declares the M rows of the state matrix.
 * rvn_gpu_enkf_pack       writes this handle's members into Xlocal[nMembers][M] (DEVICE memory) = AddToStateMatrix.
 * rvn_gpu_enkf_update     given the state rows of ALL N ensemble members Xall[N][M] (DEVICE memory; on several GPUs the
 *                         all-gather of every rank's Xlocal, members in global order) and the N x N matrix Z = HA'.inv(P).(D-HA)
 *                         (HOST memory, row-major, computed from the N x Nobs observation matrices by the host layer),
 *                         computes for the local members e = member0_global .. +nMembers-1
 *                             Xa[e][k] = X[e][k] + (1/(N-1)) * sum_i (X[i][k] - mean_i X[.][k]) * Z[i][e]
 *                         with the reference's summation order (i ascending, no FMA), writes Xa back into the caller's
 *                         Xall rows of the local members, and stores max(Xa,0) into the model state = UpdateFromStateMatrix. */
int rvn_gpu_enkf_configure(rvn_gpu_model *m, const rvn_enkf_row *rows, int64_t nrows);
int rvn_gpu_enkf_pack(rvn_gpu_model *m, double *Xlocal_dev);
int rvn_gpu_enkf_update(rvn_gpu_model *m, double *Xall_dev, const double *Z_host, int N, int member0_global);
/* the same two steps with HOST matrices (single-GPU callers, tests): X[nMembers][M] out / Xall[N][M] in-out */
int rvn_gpu_enkf_pack_host(rvn_gpu_model *m, double *Xlocal_host);
int rvn_gpu_enkf_update_host(rvn_gpu_model *m, double *Xall_host, const double *Z_host, int N, int member0_global);
/* CUDA-event time [ms] of the last rvn_gpu_enkf_update's update kernel */
int rvn_gpu_enkf_last_ms(rvn_gpu_model *m, double *update_ms);

/* raw device pointer of the state for zero-copy callers that already live on the GPU.  The state is tiled: element (member e, state
 * variable i, HRU k) lies at ((e * nTiles + k / 128) * NS + i) * 128 + k % 128 doubles, nTiles = ceil(nHRU / 128); member_stride = NS * 128 *
 * nTiles, sv_stride = 128 (the stride between state variables inside a tile) */
int rvn_gpu_device_state_ptr(rvn_gpu_model *m, void **dptr, int64_t *member_stride, int64_t *sv_stride);
/* CUDA-event time (ms) spent in kernels of the last rvn_gpu_run, and number of kernel launches it made */
int rvn_gpu_last_run_stats(rvn_gpu_model *m, double *sweep_ms, double *total_ms, int64_t *launches);

/* which HRU-sweep kernel serves this model: "static:<PROGRAM>" (compile-time specialisation of a known process
 * program, csrc/rvn_static_programs.cuh) or "interpreter" (any supported process list) */
int rvn_gpu_kernel_variant(rvn_gpu_model *m, char *buf, int buflen);
/* build tool (no device needed): text of the compile-time program struct for `desc`'s process list, as written into
 * csrc/rvn_static_programs.cuh by tools/gen_static_programs.py */
int rvn_gpu_emit_static_program(const rvn_model_desc *desc, const char *name, char *buf, int64_t buflen);

/* ---- model diagnostics on the device (reference CalculateDiagnostic, src/Diagnostics.cpp, as called by CModel::GetObjFuncVal,
 * src/StandardOutput.cpp:3282-3322): for every member of the range the diagnostic of the recorded hydrograph of subbasin `basin`
 * against an observation series, so that a calibration trial never copies its hydrograph back.  The modelled series is the one the
 * reference compares with a HYDROGRAPH observation: mod[0] = qout_initial[member], mod[nn] = 0.5 * (Q[nn-1] + Q[nn-2]) * dt / dt of
 * the recorded outflows (CModel::UpdateDiagnostics, src/Model.cpp:2929-2936); obs[nobs] (host) is the observation series sampled on
 * the model step (RVN_BLANK where missing); samples nnstart <= nn < nnend count.  Sums run in the reference's order (one thread per
 * member, sequential over time), so the value is the host layer's bit for bit.  type: 0 NASH_SUTCLIFFE, 1 RMSE, 2 KLING_GUPTA.
 * out[nmembers] (host); -1e99 where the reference returns -ALMOST_INF. */
#define RVN_BLANK (-1.2345)
int rvn_gpu_diagnostic(rvn_gpu_model *m, int type, int basin, const double *obs, int nobs, int nnstart, int nnend,
                       const double *qout_initial, int member0, int nmembers, double *out);

/* verification hook (tests/test_glibc_math.py): evaluate the device exp (fn 0) / log (1) / pow (2) / expm1 (3) / tanh (4) the kernels use
 * (csrc/rvn_glibc_math.h - restatements of the host libm the reference links, src/ calls of exp/log/pow/tanh) on n host
 * arguments x[, y] and return the results in out[n]; these must equal the host libm's bit for bit */
int rvn_gpu_math_eval(int device, int fn, int64_t n, const double *x, const double *y, double *out);

#ifdef __cplusplus
}
#endif
#endif /* RVN_GPU_H */
