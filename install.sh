#!/bin/bash
# Launcher of the B200 ray tracer.  Same command line as the reference's install.sh
# (flags -n/-f/-d/-m/-h, reference install.sh:85-102) and the same directory choreography:
# compile in src/, park objects in build/, put the program in bin/ and start it FROM bin/ so
# that it finds ../res/<settings> and writes ../data/<folder>/.
#
#   -n, --threads N   the reference: OpenMP threads.  Here: how many GPUs share the two ray
#                     loops (1 = one device, the default 32 = every visible device).
#   -f, --file FILE   settings file inside res/ (default settings.params)
#   -d, --debug       link against the assert-instrumented CUDA library (make DEBUG=1)
#   -m, --make        compile only (with -pedantic), do not run
#   -h, --help        this text
set -e
here="$(cd "$(dirname "$0")" && pwd)"

usage() {
  sed -n '2,14p' "$0" | sed 's/^# \{0,1\}//'
}

gpus=32
settings="settings.params"
mode=release
while [ $# -gt 0 ]; do
  case "$1" in
    -n|--threads) gpus="$2"; shift ;;
    -f|--file)    settings="$2"; shift ;;
    -d|--debug)   mode=debug ;;
    -m|--make)    mode=compile-only ;;
    -h|--help)    usage; exit 0 ;;
    *)            echo "install.sh: unknown option $1" >&2; usage >&2; exit 64 ;;
  esac
  shift
done

mkdir -p "$here/build" "$here/bin" "$here/data"

# --- compile (the GPU count is a run-time matter: one build serves every -n) ---------------
case "$mode" in
  debug)        target=debug ;;
  compile-only) target=build ;;
  *)            if [ "$gpus" = 1 ]; then target=all; else target=mp; fi ;;
esac
make -C "$here/src" clean
make -C "$here/src" "$target"
find "$here/src" -maxdepth 1 -name '*.o' -exec mv {} "$here/build/" \;
[ "$mode" = compile-only ] && exit 0

mv "$here/src/raytrace" "$here/bin/raytrace"
printf '\n*****Install complete*****\n\n'

# --- the unmodified reference, where a Fortran compiler and its tree exist (neither does in the build
# image): built next to the oracle so that `bench.py --impl reference` / cpu_baseline time the real thing
ref="${ORT_REFERENCE_DIR:-/root/reference}"
if command -v gfortran >/dev/null 2>&1 && [ -d "$ref/src" ]; then
  make -C "$here/oracle" ref REF="$ref" && echo "reference binary: oracle/_ref/raytrace (bench.py --impl reference times it)"
fi

# --- run, from bin/ like the reference -------------------------------------------------------
if [ "$gpus" != 32 ]; then
  export ORT_NUM_GPUS="$gpus"      # the reference exports OMP_NUM_THREADS here
  # (more than the visible devices: raytrace says so and uses what there is)
fi
cd "$here/bin"
exec ./raytrace "$settings"
