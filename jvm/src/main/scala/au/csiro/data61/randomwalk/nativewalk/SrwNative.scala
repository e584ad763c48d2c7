package au.csiro.data61.randomwalk.nativewalk

/** JNI surface over libsrw (include/srw.h).  NOT compiled in this repository (no JVM in the build image). */
object SrwNative {
  System.loadLibrary("srwjni") // libsrwjni.so links libsrw.so

  /** Main.main / runJob for `--cmd randomwalk` (Main.scala:18-27, 109-127) inside libsrw: parse, load, walk, save.
    * `argv` is exactly what CommandParser was given.  Returns 0; throws RuntimeException(srw_last_error()) otherwise. */
  @native def runRandomWalk(argv: Array[String]): Int

  /** In-memory variant: edge arrays in (the driver after `collect()`), walks out, one Array[Int] per path in
    * (round, ascending vertex id) order.  `weights` / `partitionIds` may be null.  sampler: 0 alias, 1 exact
    * (bit-identical to RandomSample.scala under the counter-based draw), 2 alias-fold (default). */
  @native def walk(src: Array[Int], dst: Array[Int], weights: Array[Float], partitionIds: Array[Int], directed: Boolean,
                   walkLength: Int, numWalks: Int, p: Double, q: Double, seed: Long, sampler: Int): Array[Array[Int]]

  /** srw_device_count(): 0 means the native path cannot run and the caller should keep the Spark implementation. */
  @native def deviceCount(): Int
}
