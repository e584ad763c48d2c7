package au.csiro.data61.randomwalk.nativewalk

import au.csiro.data61.randomwalk.common.{Params, Property}
import org.apache.spark.SparkContext
import org.apache.spark.rdd.RDD

/** Drop-in body for Main.doRandomWalk (Main.scala:53-62).  NOT compiled in this repository (no JVM in the build image).
  *
  * The reference builds UniformRandomWalk / VCutRandomWalk, calls execute() and save().  Here the whole of
  * loadGraph + randomWalk + save runs inside libsrw on the GPU; the option names, defaults and the
  * `<output>/path/part-NNNNN` files are the reference's, so everything downstream (the word2vec stage of
  * `--cmd node2vec`, Main.scala:36-44) keeps working on the files it already reads.
  */
object NativeRandomWalk {

  /** The options CommandParser accepted (CommandParser.scala:34-90), as an argv for srw_main.  `gpus` > 1 shards the graph over
    * the GPUs of the box; with `--partitioned true` (VCutRandomWalk, Main.scala:54-57) the partition-id column of the edge file is
    * then the shard map: a vertex lives on GraphMap.getPartition(v) mod gpus (VCutRandomWalk.scala:121-134). */
  def toArgv(param: Params, gpus: Int = 1): Array[String] = Array(
    "--cmd", "randomwalk",
    "--input", param.input,
    "--output", param.output,
    "--walkLength", param.walkLength.toString,
    "--numWalks", param.numWalks.toString,
    "--p", param.p.toString,
    "--q", param.q.toString,
    "--weighted", param.weighted.toString,
    "--directed", param.directed.toString,
    "--partitioned", param.partitioned.toString,
    "--rddPartitions", param.rddPartitions.toString,
    "--singleOutput", param.singleOutput.toString,
    "--gpus", gpus.toString)

  def run(context: SparkContext, param: Params, gpus: Int = 1): RDD[Array[Int]] = {
    require(SrwNative.deviceCount() >= gpus, s"libsrw needs $gpus CUDA device(s)")
    SrwNative.runRandomWalk(toArgv(param, gpus))
    // what RandomWalk.save wrote (RandomWalk.scala:234-241), back as the RDD the embedding stage expects
    context.textFile(s"${param.output}/${Property.pathSuffix}").map(_.split("\t").map(_.toInt))
  }
}
