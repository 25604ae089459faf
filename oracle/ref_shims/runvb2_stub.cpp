// TEST INFRASTRUCTURE ONLY (oracle/): the pop+con stage (VerifyBamID2, needs
// htslib + Eigen) is outside the align hot path; FASTQuick_ref links this stub.
int runVB2(int, char **) { return 1; }
