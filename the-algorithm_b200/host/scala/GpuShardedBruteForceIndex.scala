package com.twitter.ann.brute_force

// ShardedAppendable + ComposedQueryable (common/ShardApi.scala:34-48, 58-87) over the GPUs of one box, in ONE JVM: the
// process model the reference's sharding has (Future.collect over in-process sub-indices).  One native handle
// (ann_sharded_*, include/b200ann.h) owns one shard per device; an appended batch is cut into contiguous parts that travel
// to their devices in parallel, and a query batch runs the three-phase sharded query on every device at once (seed bounds
// -> tensor-core filter -> exact rescoring of the device's share of the global survivors), ordered across devices by CUDA
// events; each device merges and returns its slice of the batch.  The answer is bit for bit what one index over all rows
// returns (ids unique across shards), ties by id where the reference's stable sort keeps shard order (ShardApi.scala:80-84).
//
// NOT COMPILED IN THIS REPOSITORY'S CI (no JVM in the image); same status as GpuBruteForceIndex.scala.  The native handle is
// exercised by tests/test_gpu_round2.py (ctypes) and tests/cpp_host_check.cpp (C++).

import com.twitter.ann.common.Appendable
import com.twitter.ann.common.Distance
import com.twitter.ann.common.EmbeddingType._
import com.twitter.ann.common.EntityEmbedding
import com.twitter.ann.common.Metric
import com.twitter.ann.common.NeighborWithDistance
import com.twitter.ann.common.Queryable
import com.twitter.ann.common.{Cosine, InnerProduct, L2}
import com.twitter.util.Future
import com.twitter.util.FuturePool
import java.nio.{ByteBuffer, ByteOrder}

object GpuShardedBruteForceIndex {
  private def ordinal(metric: Metric[_]): Int = metric match { // thrift DistanceMetric, ann_common.thrift:16-19
    case L2 => 0
    case Cosine => 1
    case InnerProduct => 2
    case other => throw new IllegalArgumentException(s"metric $other is not a dense-vector metric")
  }

  /** `devices`: CUDA ordinals, one shard each (what `shards` is to ShardedAppendable, ShardApi.scala:34-37). */
  def apply[D <: Distance[D]](
    metric: Metric[D],
    futurePool: FuturePool,
    dimension: Int,
    devices: Seq[Int],
    initialEmbeddings: Iterator[EntityEmbedding[Long]] = Iterator()
  ): GpuShardedBruteForceIndex[D] = {
    val index = new GpuShardedBruteForceIndex[D](metric, futurePool, dimension, devices.toArray, 0L)
    initialEmbeddings.grouped(1 << 18).foreach(batch => index.appendBatch(batch))
    index
  }

  /** ComposedQueryableDeserialization.fromDirectory (ShardedSerialization.scala:49-66): reads every `shard_<i>/` directory
   * (written by this class or by the reference's ShardedSerialization) and re-deals the rows over `devices`. */
  def fromDirectory[D <: Distance[D]](
    directory: String,
    metric: Metric[D],
    futurePool: FuturePool,
    devices: Seq[Int]
  ): GpuShardedBruteForceIndex[D] = {
    val h = B200AnnNative.shardedLoadDirectory(ordinal(metric), 0, 0, directory, 0, devices.toArray)
    if (h == 0L) throw new RuntimeException(B200AnnNative.lastError())
    new GpuShardedBruteForceIndex[D](metric, futurePool, -1, devices.toArray, h)
  }
}

class GpuShardedBruteForceIndex[D <: Distance[D]] private (
  metric: Metric[D],
  futurePool: FuturePool,
  dimension: Int,
  devices: Array[Int],
  loaded: Long)
    extends Appendable[Long, BruteForceRuntimeParams.type, D]
    with Queryable[Long, BruteForceRuntimeParams.type, D]
    with AutoCloseable {

  private[this] var handle: Long =
    if (loaded != 0L) loaded
    else {
      val h = B200AnnNative.shardedCreate(GpuShardedBruteForceIndex.ordinal(metric), dimension, 0L, 0, devices)
      if (h == 0L) throw new RuntimeException(B200AnnNative.lastError())
      h
    }

  private[this] def check(rc: Int): Unit =
    if (rc != 0) throw new RuntimeException(s"b200ann error $rc: ${B200AnnNative.lastError()}")

  private[this] def direct(bytes: Int): ByteBuffer = ByteBuffer.allocateDirect(bytes).order(ByteOrder.nativeOrder())

  /** ShardedAppendable.append for a whole batch: part s of the batch goes to shard s (deterministic and balanced where the
   * reference draws Random.nextInt(shards) per row, ShardApi.scala:21-25). */
  def appendBatch(batch: Seq[EntityEmbedding[Long]]): Unit = synchronized {
    if (batch.nonEmpty) {
      val d = batch.head.embedding.length
      val ids = direct(batch.size * 8)
      val rows = direct(batch.size * d * 4)
      batch.foreach { e =>
        ids.putLong(e.id)
        var i = 0
        while (i < d) { rows.putFloat(e.embedding(i)); i += 1 }
      }
      check(B200AnnNative.shardedAppendBatch(handle, ids, rows, batch.size.toLong))
    }
  }

  override def append(embedding: EntityEmbedding[Long]): Future[Unit] = futurePool { appendBatch(Seq(embedding)) }

  override def toQueryable: Queryable[Long, BruteForceRuntimeParams.type, D] = this

  /** ComposedQueryable.queryWithDistance (ShardApi.scala:72-86) for b queries in one native call. */
  def batchQueryWithDistance(
    embeddings: Seq[EmbeddingVector],
    numOfNeighbours: Int
  ): Seq[List[NeighborWithDistance[Long, D]]] = {
    if (numOfNeighbours <= 0 || embeddings.isEmpty) return embeddings.map(_ => Nil)
    val b = embeddings.size
    val d = embeddings.head.length
    val q = direct(b * d * 4)
    embeddings.foreach { e => var i = 0; while (i < d) { q.putFloat(e(i)); i += 1 } }
    val outIds = direct(b * numOfNeighbours * 8)
    val outDist = direct(b * numOfNeighbours * 4)
    val outCount = direct(b * 4)
    check(B200AnnNative.shardedQueryBatch(handle, q, b, d, numOfNeighbours, outIds, outDist, outCount))
    (0 until b).map { qi =>
      (0 until outCount.getInt(qi * 4)).map { j =>
        NeighborWithDistance(
          outIds.getLong((qi * numOfNeighbours + j) * 8),
          metric.fromAbsoluteDistance(outDist.getFloat((qi * numOfNeighbours + j) * 4)))
      }.toList
    }
  }

  override def queryWithDistance(
    embedding: EmbeddingVector,
    numOfNeighbours: Int,
    runtimeParams: BruteForceRuntimeParams.type
  ): Future[List[NeighborWithDistance[Long, D]]] =
    futurePool { batchQueryWithDistance(Seq(embedding), numOfNeighbours).head }

  override def query(
    embedding: EmbeddingVector,
    numOfNeighbours: Int,
    runtimeParams: BruteForceRuntimeParams.type
  ): Future[List[Long]] =
    queryWithDistance(embedding, numOfNeighbours, runtimeParams).map(_.map(_.neighbor))

  /** ShardedSerialization.toDirectory (ShardedSerialization.scala:28-38): `shard_<i>/BruteForceFileData`, thrift stream. */
  def toDirectory(directory: String): Unit = synchronized { check(B200AnnNative.shardedSaveDirectory(handle, directory, 1, 0)) }

  def size: Long = synchronized { B200AnnNative.shardedSize(handle) }

  override def close(): Unit = synchronized {
    if (handle != 0L) { B200AnnNative.shardedDestroy(handle); handle = 0L }
  }
}
