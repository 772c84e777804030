// Generates tests/golden/javafastpfor_0.1.10.json: real output of the library the reference's PFORCodecInt delegates to
// (core/src/main/scala/immutabledb/codec/PFORCodec.scala:7,15-18; project/Dependencies.scala:4).  This image has no JVM;
// anyone who has one turns the codec's "parity unpinned" status into pinned with:
//
//   scala -cp JavaFastPFOR-0.1.10.jar tools/gen_javafastpfor_goldens.scala > tests/golden/javafastpfor_0.1.10.json
//   python -m pytest tests/test_codec_pinning.py -q
//
import me.lemire.integercompression.differential.IntegratedIntCompressor

object GenGoldens extends App {
  val rnd = new scala.util.Random(42)
  def sorted(n: Int, step: Int, start: Int) = Array.iterate(start, n)(_ + rnd.nextInt(step + 1))
  val cases: Seq[(String, Array[Int])] = Seq(
    "empty" -> Array[Int](),
    "single300" -> Array(300),
    "seq32" -> Array.range(0, 32),
    "seq1024_from_5e8" -> Array.range(500000000, 500001024),
    "worksheet15" -> Array(1, 2, 3, 5, 8, 13, 21, 34, 55, 89, 144, 233, 377, 610, 987),
    "seq128_plus_tail7" -> Array.range(1000, 1135),
    "three_superblocks_two_miniblocks_tail" -> sorted(128 * 3 + 64 + 9, 300, 17),
    "steps_b2" -> sorted(256, 3, -2000000000),
    "steps_b13" -> sorted(1024, 8191, 0),
    "constant" -> Array.fill(160)(123456789),
    "negative_delta_raw_miniblock" -> (Array(5, 3) ++ Array.fill(30)(3)),
    "unsorted" -> Array.fill(200)(rnd.nextInt()),
    "extremes" -> Array(Int.MinValue, Int.MaxValue, 0, -1, 1, Int.MinValue, Int.MaxValue) ,
    "wraparound" -> Array.iterate(Int.MaxValue - 40, 96)(_ + 1)
  )
  val iic = new IntegratedIntCompressor()
  val body = cases.map { case (name, in) =>
    val out = iic.compress(in)
    s"""{"name":"$name","input":[${in.mkString(",")}],"compressed":[${out.mkString(",")}]}"""
  }
  println(s"""{"library":"JavaFastPFOR 0.1.10","class":"me.lemire.integercompression.differential.IntegratedIntCompressor","cases":[${body.mkString(",")}]}""")
}
