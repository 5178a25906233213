#define SVT_AV1_CVS_VERSION "v0.8.6-ref-oracle"
