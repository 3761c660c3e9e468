#!/bin/bash
# BASELINE config C5 as the reference runs it: the UNCHANGED Victoria Park driver with 4 000 particles on the head of the dataset
# (2 500 sensor messages = 254 lidar updates): on the reference's filter header (OpenMP, all host threads), on the drop-in header +
# librfsb200 (fp32 device build), on the drop-in with device-side particle propagation (RFSB200_DEVICE_PROPAGATE=1), and on the
# drop-in with the candidate lists of addBirthGaussians kept on the host (RFSB200_HOST_BIRTHS=1, the round-1 arrangement).
# Prints wall time and the filter's own timing table of each.
R=/root/repo/oracle/_ref
sed -e 's|<nParticles>100</nParticles>|<nParticles>4000</nParticles>|' -e 's|<effNParticle>50.0</effNParticle>|<effNParticle>2000.0</effNParticle>|' \
    -e 's|<logResultsToFile>1</logResultsToFile>|<logResultsToFile>0</logResultsToFile>|' $R/rbphdslam_VictoriaPark.xml > /tmp/vp4000.xml
for B in b200dev b200 b200hostbirths ref; do
  mkdir -p /tmp/c5_$B && cd /tmp/c5_$B && ln -sfn $R/vpdata vpdata
  BIN=$B; DEV=0; HB=0
  if [ $B = b200dev ]; then BIN=b200; DEV=1; fi
  if [ $B = b200hostbirths ]; then BIN=b200; HB=1; fi
  S=$(date +%s.%N)
  RFSB200_HOST_BIRTHS=$HB RFSB200_DEVICE_PROPAGATE=$DEV RFSB200_GM_CAPACITY=192 $R/rbphdslam_VictoriaPark_$BIN -c /tmp/vp4000.xml -s 1 > run.log 2>&1
  E=$(date +%s.%N)
  echo "== $B: wall $(python3 -c "print('%.2f' % ($E - $S))") s, nproc $(nproc)"
  grep -E "Prediction|Map Update  |Resampling|Total" run.log | head -4
done
