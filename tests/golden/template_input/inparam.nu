# inparam.nu of the reference's template/input with the comment blocks removed: keyword / value lines only
# (fixture for tests/main_case.py; the golden vectors were made from the unchanged files, oracle/make_golden_main.py)
NU_TYPE                                     constant
NU_CONST                                    2
NU_EMP_REF                                  24
NU_EMP_MIN                                  8
NU_EMP_SCALE_AXIS                           true
NU_EMP_POW_AXIS                             1.0
NU_EMP_SCALE_THETA                          true
NU_EMP_POW_THETA                            3.0
NU_EMP_FACTOR_PI                            5.0
NU_EMP_THETA_START                          45.0
NU_EMP_SCALE_DEPTH                          true
NU_EMP_FACTOR_SURF                          2.0
NU_EMP_DEPTH_START                          200.0
NU_EMP_DEPTH_END                            300.0
NU_WISDOM_LEARN                             false
NU_WISDOM_LEARN_EPSILON                     1e-3
NU_WISDOM_LEARN_INTERVAL                    5
NU_WISDOM_LEARN_OUTPUT                      name.nu_wisdom.nc
NU_WISDOM_REUSE_INPUT                       name.nu_wisdom.nc
NU_WISDOM_REUSE_FACTOR                      1.0
NU_USER_PARAMETER_LIST                      -1.2345
