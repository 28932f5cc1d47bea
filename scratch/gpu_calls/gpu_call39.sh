for l in tomosipo_b200/libtsproj.so scratch/lib_s3.so scratch/lib_s6.so scratch/lib_s8.so; do echo $l; TSPROJ_LIB=$PWD/$l python scratch/prof_step.py 512 720 3 2>&1 | tail -3 | head -2; done
