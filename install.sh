#!/bin/bash
# Build and run the B200 ray tracer -- same interface as the reference's install.sh
# (reference install.sh:85-102): builds in src/, moves the binary to bin/, runs it from bin/.

function showhelp
{
  echo 'Usage: ./install.sh [option] [option]...'
  echo 'Compile and run raytrace (B200 build).'
  echo
  echo '   -h, --help            Shows this dialog.'
  echo '   -n, --threads         Number of GPUs that share the ray loops (the reference: OpenMP'
  echo '                         threads).  1 = a single device; default 32 = all visible devices.'
  echo '   -f, --file            Settings file in res/ (default settings.params).'
  echo '   -d, --debug           Device debug build (-G) of the CUDA library.'
  echo '   -m, --make            Compile only, with warnings enabled.'
}

function makebuild
{
  if [ "$debug" = 1 ]; then
    make clean && make debug
  elif [ "$make" = 1 ]; then
    make clean && make build
  else
    if [ "$NUM_THREADS" = 1 ]; then
      make clean && make
    else
      make clean && make mp
    fi
  fi
}

function createdirs
{
  if [ ! -d "build" ]; then mkdir "build"; fi
  cd build; ndirec="$(pwd)"; cd ..
  if [ ! -d "bin" ]; then mkdir "bin"; fi
  cd bin; bdirc="$(pwd)"; cd ..
  if [ ! -d "data" ]; then mkdir "data"; fi
  cd src
}

function run
{
  for i in *; do
    if [ "${i}" != "${i%.o}" ]; then mv "${i}" "$ndirec"; fi
  done
  if [ "$make" = "1" ]; then exit 0; fi
  mv raytrace "$bdirc" && echo " " && echo "*****Install complete*****" && echo " "
  cd ../bin
  ./raytrace $file
}

#defaults
NUM_THREADS=32
debug=0
help=0
make=0
file="settings.params"
set -e
cd "$(dirname "$0")"

createdirs

while [ "$1" != "" ]; do
    case $1 in
        -n | --threads )        NUM_THREADS=$2
                                ;;
        -h | --help )           showhelp
                                exit
                                ;;
        -m | --make )           make=1
                                makebuild
                                exit
                                ;;
        -d | --debug )          debug=1
                                ;;
        -f | --file )           file=$2
                                ;;
    esac
    shift
done

makebuild
# the reference exports OMP_NUM_THREADS here; the GPU build caps the device count instead
if [ "$NUM_THREADS" != "32" ]; then
  export ORT_NUM_GPUS=$NUM_THREADS
fi
run
