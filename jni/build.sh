#!/bin/bash
# Builds libvce_jni.so (and compiles the Java side against the reference's classes) where a JDK is present.
# Usage: jni/build.sh [path-to-rtg-tools-classes-or-jar]
set -e
HERE="$(cd "$(dirname "$0")" && pwd)"
ROOT="$(dirname "$HERE")"
JAVAC="$(command -v javac || true)"
if [ -z "$JAVAC" ]; then
  echo "jni/build.sh: no javac on PATH -- nothing built (the C side needs jni.h from a JDK)"
  exit 0
fi
JAVA_HOME_GUESS="$(dirname "$(dirname "$(readlink -f "$JAVAC")")")"
JH="${JAVA_HOME:-$JAVA_HOME_GUESS}"
if [ ! -f "$JH/include/jni.h" ]; then
  echo "jni/build.sh: $JH/include/jni.h not found -- nothing built"
  exit 0
fi
mkdir -p "$HERE/build"
gcc -O2 -fPIC -shared -Wall -I"$JH/include" -I"$JH/include/linux" -o "$HERE/build/libvce_jni.so" "$HERE/vce_jni.c" \
    -L"$ROOT/rtg_tools_b200/csrc" -lvce -Wl,-rpath,"$ROOT/rtg_tools_b200/csrc"
echo "built $HERE/build/libvce_jni.so"
if [ -n "$1" ]; then
  # the Java classes live in the reference's package (they use its package-private types): compile against its classes
  "$JAVAC" -cp "$1" -d "$HERE/build/classes" "$HERE"/*.java
  echo "compiled the Java side into $HERE/build/classes"
fi
